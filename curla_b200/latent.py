"""Batched latent / Q-value extraction: the per-observation loop of
plot_tsne/latent_data.py:63-104 (BASELINE.json configs[4]).

The reference walks a replay buffer one observation at a time: center-crop on the host
(augmentor.evaluation_augmentation), H2D copy, `agent.actor.encoder(gpu_obs)`, a sampled action
(`agent.sample_action`, a second encoder forward), `agent.critic(gpu_obs, gpu_action)` (a third)
and three D2H syncs per observation.  Here the uint8 frames go to the device once, the fused
gather+crop kernel takes the center window, and every network runs on whole batches: one actor
encoder pass, the policy head, one critic encoder pass and the batched Q1 || Q2 heads.
There is no CPU fallback (the agent's engine raises without a CUDA device).
"""
import numpy as np
import torch


def center_window(frame_hw, image_hw):
    """(top, left) of augmentations.py:37-43's center crop; (0, 0) when shapes agree."""
    return (frame_hw[0] - image_hw[0]) // 2, (frame_hw[1] - image_hw[1]) // 2


def extract(agent, obses, sample=True, noise=None):
    """obses: uint8 array / tensor (N, C, Hf, Wf) (stored frames, e.g. CustomReplayBuffer.obses).

    Returns dict(representations float32 (N, feature_dim)  -- agent.actor.encoder latents,
                 actions float32 (N, A)                     -- pi (sample=True, like sample_action) or tanh(mu),
                 q_values float32 (N,)                       -- min(Q1, Q2)(obs, action), latent_data.py:87-101)
    as numpy arrays.  `noise` (N, A) injects the policy noise (tests)."""
    frames = torch.as_tensor(np.ascontiguousarray(obses) if isinstance(obses, np.ndarray) else obses)
    if frames.dtype != torch.uint8:
        raise TypeError('extract: stored observations must be uint8 frames')
    dev = agent.device
    frames = frames.to(dev)
    top, left = center_window(tuple(frames.shape[-2:]), agent.image_shape)
    with torch.no_grad():
        z_actor = agent.encode_frames(0, frames, top, left)                      # actor.encoder (LayerNorm output)
        if noise is not None:
            noise = torch.as_tensor(noise, dtype=torch.float32, device=dev)
        mu, pi, _, _ = agent.actor_head(z_actor, compute_pi=sample, compute_log_pi=False, noise=noise)
        action = pi if sample else mu
        z_critic = agent.encode_frames(1, frames, top, left)                     # critic.encoder (own fc / ln)
        q1, q2 = agent.q_heads(1, z_critic, action)
        q = torch.minimum(q1, q2).squeeze(1)
    return dict(representations=z_actor.cpu().numpy().astype(np.float32), actions=action.cpu().numpy().astype(np.float32),
                q_values=q.cpu().numpy().astype(np.float32))
