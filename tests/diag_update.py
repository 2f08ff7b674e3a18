"""Diagnostic (not a test): print every CUDA-vs-oracle error of the update for a scenario.
usage: python tests/diag_update.py [scenario ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import scenario as S
import test_update_parity_gpu as T


def rl2(a, b):
    return T.rel_l2(a, b)


def main():
    names = sys.argv[1:] or list(S.SCENARIOS)
    for name in names:
        cfg = S.SCENARIOS[name]
        run = S.OracleRun(cfg)
        agent, rb = T.build_cuda_agent(cfg, run)
        agent2, _ = T.build_cuda_agent(cfg, run)     # forward-only probe fed with the ORACLE's weights
        st = np.random.get_state()
        L = T.NullLogger()
        for u, (step, only_cpc) in enumerate(zip(cfg['steps'], cfg['only_cpc'])):
            np.random.set_state(st)
            d, b, om = run.step()
            after = np.random.get_state()
            np.random.set_state(st)
            agent._noise_override = (run.noise[u, 0], run.noise[u, 1])
            agent.update(rb, L, step, only_cpc=only_cpc)
            torch.cuda.synchronize()
            st = after
            eng, o = agent.engine, run.agent
            print('==== %s update %d (step %d only_cpc %s) launches %d' % (name, u, step, only_cpc, eng.last_launches()))
            km = {'train/batch_reward': 'batch_reward', 'train_critic/loss': 'critic_loss', 'train_actor/loss': 'actor_loss',
                  'train_actor/entropy': 'entropy', 'train_alpha/loss': 'alpha_loss', 'train/curl_loss': 'curl_loss'}
            for rk, ok in km.items():
                if ok in om:
                    print('  %-22s cuda % .6f oracle % .6f' % (rk, L.rows[(step, rk)], om[ok]))
            fd = S.FEATURE_DIM
            if not only_cpc:
                print('  z_critic %.2e  q1 %.2e  target_q %.2e  next_action %.2e next_logpi %.2e' % (
                    rl2(eng.t['p3.z'][:, :fd], o.dbg['z_critic']), rl2(eng.t['p3.q1.out'], o.dbg['q1']),
                    rl2(eng.t['target_q'], o.dbg['target_q'][:, 0]), rl2(eng.t['next_action'], o.dbg['next_action']),
                    rl2(eng.t['next_log_pi'], o.dbg['next_log_pi'][:, 0])))
            if 'pi' in o.dbg:
                print('  pi %.2e log_pi %.2e' % (rl2(eng.t['pi'], o.dbg['pi']), rl2(eng.t['log_pi'], o.dbg['log_pi'][:, 0])))
            if 'z_a' in o.dbg:
                print('  z_a %.2e z_pos %.2e' % (rl2(eng.t['p5.z'][:, :fd], o.dbg['z_a']), rl2(eng.t['p7.z'][:, :fd], o.dbg['z_pos'])))
            gv = T.grad_views(agent)
            for tag, og in (('critic_opt', o.dbg.get('critic_grads')), ('actor_opt', o.dbg.get('actor_grads')),
                            ('cpc_opt', o.dbg.get('cpc_grads'))):
                if og is None:
                    continue
                net = 'actor.' if tag == 'actor_opt' else 'critic.'
                for k, g_ref in og.items():
                    ek = net + (k.replace('fc.weight', 'fc.weight_canon') if k == 'encoder.fc.weight' else k)
                    ours = T.to_torch_layout(eng, ek, gv[tag + '/' + ek])
                    print('  grad %-10s %-26s rel %.2e  |ref| %.2e' % (tag, k, rl2(ours, g_ref), float(g_ref.norm())))
            if 'W_grad' in o.dbg:
                print('  grad W rel %.2e' % rl2(gv['cpc_opt/W'], o.dbg['W_grad']))
            for net, osd in (('actor', o.actor), ('critic', o.critic), ('target', o.target)):
                worst = (0, None, 0)
                for k, v in osd.items():
                    if net == 'actor' and k.startswith('encoder.convs.'):
                        continue
                    ek = net + '.' + (k.replace('fc.weight', 'fc.weight_canon') if k == 'encoder.fc.weight' else k)
                    ours = T.to_torch_layout(eng, ek, eng.t[ek]).cpu()
                    err = (ours - v.detach()).abs()
                    if float(err.max()) > worst[0]:
                        worst = (float(err.max()), k, float(err.mean()))
                print('  params %-7s worst max|err| %.2e (%s) mean %.2e' % (net, worst[0], worst[1], worst[2]))
            print('  log_alpha cuda %.9f oracle %.9f' % (float(agent.log_alpha), float(o.log_alpha.detach())))
            # forward-only: CUDA encoders evaluated with the oracle's post-update weights
            sd = lambda dct: {k: v.detach() for k, v in dct.items()}
            agent2.critic.load_state_dict(sd(o.critic)); agent2.actor.load_state_dict(sd(o.actor))
            agent2.critic_target.load_state_dict(sd(o.target))
            obs_f = torch.from_numpy(b['obs']).float().cuda(); pos_f = torch.from_numpy(b['pos']).float().cuda()
            import oracle.curla_oracle as O
            with torch.no_grad():
                za_o = O.encoder_forward(o.critic, 'encoder.', obs_f.cpu()); zp_o = O.encoder_forward(o.target, 'encoder.', pos_f.cpu())
            print('  [same weights] critic.encoder(obs) %.2e  target.encoder(pos) %.2e  actor.encoder(obs) %.2e' % (
                rl2(agent2.critic.encoder(obs_f), za_o), rl2(agent2.critic_target.encoder(pos_f), zp_o),
                rl2(agent2.actor.encoder(obs_f), O.encoder_forward(o.actor, 'encoder.', obs_f.cpu()).detach())))
            # how many Adam sign decisions differ (critic segment)
            for net, osd in (('critic', o.critic),):
                tot = flips = 0
                for k, v in osd.items():
                    ek = net + '.' + (k.replace('fc.weight', 'fc.weight_canon') if k == 'encoder.fc.weight' else k)
                    ours = T.to_torch_layout(eng, ek, eng.t[ek]).cpu()
                    e = (ours - v.detach()).abs()
                    tot += e.numel(); flips += int((e > 0.5e-3).sum())
                print('  critic elements with |err| > lr/2: %d of %d (%.2f%%)' % (flips, tot, 100.0 * flips / tot))


if __name__ == '__main__':
    main()
