"""ctypes binding of libcurla_b200.so (the C ABI declared in include/curla_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, this
raises.  PyTorch is used for device memory and streams only.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libcurla_b200.so')

c_ll = C.c_longlong
c_vp = C.c_void_p


class ConvSeg(C.Structure):
    """curla_conv_seg (include/curla_b200.h): one pass of a multi-pass conv launch."""
    _fields_ = [('inp', c_vp), ('wts', c_vp), ('bias', c_vp), ('out', c_vp), ('B', C.c_int)]


class ConvStackSeg(C.Structure):
    """curla_conv_stack_seg: one pass of the fused conv-2..4 launch."""
    _fields_ = [('inp', c_vp), ('w96', c_vp), ('bias', c_vp * 3), ('out', c_vp * 3), ('B', C.c_int)]


class GatherSeg(C.Structure):
    """curla_gather_seg: one stream of a multi-stream gather launch."""
    _fields_ = [('frames', c_vp), ('h1', c_vp), ('w1', c_vp), ('out', c_vp)]


class GatherRows(C.Structure):
    """curla_gather_rows: the batch's action / reward / not_done rows gathered by the same launch."""
    _fields_ = [('actions', c_vp), ('rewards', c_vp), ('not_dones', c_vp), ('out_actions', c_vp),
                ('out_rewards', c_vp), ('out_not_dones', c_vp), ('action_dim', C.c_int)]


class AgentConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        'C', 'H', 'W', 'Hf', 'Wf', 'feature_dim', 'hidden_dim', 'action_dim', 'num_filters',
        'num_layers', 'batch', 'global_batch', 'rank', 'world', 'detach_encoder', 'pixel_sac',
        'actor_update_freq', 'critic_target_update_freq', 'cpc_update_freq')] + \
        [(n, C.c_double) for n in (
            'discount', 'critic_tau', 'encoder_tau', 'actor_lr', 'actor_beta', 'critic_lr',
            'critic_beta', 'alpha_lr', 'alpha_beta', 'encoder_lr', 'log_std_min', 'log_std_max',
            'target_entropy')]


class UpdateArgs(C.Structure):
    _fields_ = [(n, c_vp) for n in (
        'obses', 'next_obses', 'actions', 'rewards', 'not_dones', 'idxs', 'h1_obs', 'w1_obs',
        'h1_next', 'w1_next', 'h1_pos', 'w1_pos', 'obs_f32', 'next_f32', 'pos_f32',
        'noise_next', 'noise_cur')] + \
        [('seed', C.c_ulonglong), ('offset', C.c_ulonglong), ('step', C.c_int),
         ('only_cpc', C.c_int), ('pos_is_obs', C.c_int), ('phases', C.c_int)]


# name -> (restype, argtypes); mirrors include/curla_b200.h one to one
_i, _f, _d = C.c_int, C.c_float, C.c_double
SIGNATURES = {
    'curla_last_error': (C.c_char_p, []),
    'curla_version': (_i, []),
    'curla_gather_crop_f32': (_i, [c_vp, _i, _i, _i, c_vp, c_vp, c_vp, _i, _i, _i, c_vp, c_vp]),
    'curla_gather_crop_s2d': (_i, [c_vp, _i, _i, _i, c_vp, c_vp, c_vp, _i, _i, _i, _i, c_ll, c_vp, c_vp]),
    'curla_gather_crop_s2d_multi': (_i, [c_vp, _i, c_vp, _i, _i, _i, _i, _i, _i, _i, c_ll, c_vp, c_vp]),
    'curla_f32_to_s2d': (_i, [c_vp, _i, _i, _i, _i, _i, c_ll, c_vp, c_vp]),
    'curla_scatter_transition': (_i, [c_vp, _i, c_ll, c_vp, c_vp, c_vp, c_vp]),
    'curla_replay_add': (_i, [c_vp, c_vp, c_ll, c_vp, c_vp, _i, c_ll, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'curla_gather_rows_f32': (_i, [c_vp, c_vp, _i, _i, c_vp, c_vp]),
    'curla_color_jiggle': (_i, [c_vp, _i, _i, _i, c_vp, C.c_ulonglong, C.c_ulonglong, _f, _f, _f, _f, _i, c_vp, c_vp]),
    'curla_noisy_cover': (_i, [c_vp, _i, _i, _i, _i, _i, c_vp, _f, c_vp, C.c_ulonglong, C.c_ulonglong, c_vp]),
    'curla_conv_pad_rows': (_i, [_i]),
    'curla_conv_fwd': (_i, [c_vp, c_ll, c_vp, c_vp, _f, c_vp, c_ll, _i, _i, _i, _i, _i, _i, c_vp]),
    'curla_conv_dgrad': (_i, [c_vp, c_ll, c_vp, c_vp, c_vp, c_ll, _i, _i, _i, _i, _i, c_vp]),
    'curla_conv_stack_fits': (_i, [_i, _i, c_vp, c_vp]),
    'curla_conv_stack_fwd': (_i, [c_vp, _i, c_ll, _i, _i, c_vp, c_vp, c_vp]),
    'curla_agent_set_keep_acts': (_i, [c_vp, _i]),
    'curla_agent_set_mailbox': (_i, [c_vp, c_vp]),
    'curla_publish_metrics': (_i, [c_vp, c_vp, C.c_uint, c_vp, c_vp]),
    'curla_conv_fwd_multi': (_i, [c_vp, _i, c_ll, _f, c_ll, _i, _i, _i, _i, _i, c_vp]),
    'curla_conv_debug_read': (_i, [c_vp, _i]),
    'curla_gemm_tc_debug_read': (_i, [c_vp]),
    'curla_gemm_tc_stamps_read': (_i, [c_vp, _i]),
    'curla_conv_wgrad_workspace_floats': (c_ll, [_i]),
    'curla_conv_wgrad': (_i, [c_vp, c_ll, c_vp, c_ll, c_vp, c_vp, c_vp, _f, _i, _i, _i, _i, _i, _i, _i, c_vp]),
    'curla_conv_wgrad_partial': (_i, [c_vp, c_ll, c_vp, c_ll, c_vp, _i, _i, _i, _i, _i, _i, c_vp, c_vp]),
    'curla_conv_wgrad_reduce_multi': (_i, [_i, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'curla_gemm_bf16': (_i, [c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, _i, _i, _i, _i, _i, _i, c_vp, _i, c_vp,
                             c_ll, _i, c_ll, _f, c_vp]),
    'curla_gemm_bf16_seg': (_i, [c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, _i, _i, _i, _i, _i, _i, c_vp, _i, c_vp,
                                 c_ll, _i, c_ll, _f, _i, c_ll, _i, c_vp]),
    'curla_gemm_effective_splits': (_i, [_i, _i]),
    'curla_gemm_bf16_batched': (_i, [c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, _i, _i, _i, _i, _i, _i, c_vp, _i, c_vp,
                                     c_ll, _f, _i, c_ll, c_ll, c_ll, c_ll, c_ll, c_vp]),
    'curla_ln_fwd_x': (_i, [c_vp, _i, c_ll, c_vp, c_vp, c_vp, _i, _i, _i, c_vp, c_vp, c_vp, _i, c_vp, c_vp]),
    'curla_head_fwd_batched': (_i, [c_vp, _i, c_vp, c_vp, _i, _i, _i, c_vp, _i, c_ll, c_ll, c_ll, c_vp]),
    'curla_head_bwd_batched': (_i, [c_vp, c_vp, c_vp, _i, _i, _i, c_vp, _i, c_ll, c_ll, c_ll, c_ll, c_vp]),
    'curla_head_wgrad_batched': (_i, [c_vp, c_vp, _i, _i, _i, c_vp, c_vp, _i, c_ll, c_ll, c_ll, c_vp]),
    'curla_colsum_bf16_batched': (_i, [c_vp, _i, _i, c_vp, _i, c_ll, c_ll, c_vp]),
    'curla_ln_fwd': (_i, [c_vp, _i, c_ll, c_vp, c_vp, c_vp, _i, _i, _i, c_vp, c_vp, c_vp]),
    'curla_ln_bwd': (_i, [c_vp, c_vp, c_vp, c_vp, _i, _i, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'curla_pack_x': (_i, [c_vp, c_vp, _i, _i, _i, c_vp, c_vp]),
    'curla_head_fwd': (_i, [c_vp, _i, c_vp, c_vp, _i, _i, _i, c_vp, c_vp]),
    'curla_head_bwd': (_i, [c_vp, c_vp, c_vp, _i, _i, _i, c_vp, c_vp]),
    'curla_head_wgrad': (_i, [c_vp, c_vp, _i, _i, _i, c_vp, c_vp, c_vp]),
    'curla_colsum_bf16': (_i, [c_vp, _i, _i, c_vp, c_vp]),
    'curla_policy_fwd': (_i, [c_vp, c_vp, C.c_ulonglong, C.c_ulonglong, _i, _i, _f, _f, _i, _i, c_vp, c_vp,
                              c_vp, c_vp, c_vp, c_vp]),
    'curla_policy_fwd_rows': (_i, [c_vp, c_vp, C.c_ulonglong, C.c_ulonglong, _i, _i, _i, _f, _f, _i, _i, c_vp, c_vp,
                              c_vp, c_vp, c_vp, c_vp]),
    'curla_policy_fwd_rows_dyn': (_i, [c_vp, c_vp, C.c_ulonglong, C.c_ulonglong, c_vp, _i, _i, _i, _f, _f, _i, _i, c_vp, c_vp,
                                  c_vp, c_vp, c_vp, c_vp]),
    'curla_policy_bwd': (_i, [c_vp, c_vp, _i, c_vp, c_vp, c_vp, c_vp, c_vp, _i, _i, _f, _f, c_vp, c_vp]),
    'curla_critic_loss': (_i, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, _f, c_vp, c_vp, _i, _f, c_vp, c_vp, c_vp,
                               c_vp, c_vp]),
    'curla_actor_loss': (_i, [c_vp, c_vp, c_vp, c_vp, _i, _i, c_vp, _f, _f, c_vp, c_vp, c_vp, c_vp, c_vp,
                              c_vp]),
    'curla_add2': (_i, [c_vp, c_vp, c_ll, c_vp, c_vp]),
    'curla_curl_workspace_floats': (c_ll, [_i, _i]),
    'curla_curl_fwd_bwd': (_i, [c_vp, c_vp, c_vp, _i, _i, _i, _i, _f, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'curla_adam_f32': (_i, [c_vp, c_vp, c_vp, c_vp, c_ll, c_ll, _d, _d, _d, _d, _i, c_vp, c_vp]),
    'curla_set_dev_state': (_i, [c_vp, _i, _i, _i, _i, C.c_ulonglong, c_vp]),
    'curla_adam_f64_scalar': (_i, [c_vp, c_vp, c_vp, _d, _d, _d, _d, _i, c_vp, c_vp]),
    'curla_ema_f32': (_i, [c_vp, c_vp, c_ll, c_ll, _d, _d, c_vp]),
    'curla_pack_shadows': (_i, [c_vp, c_vp, c_vp, _i, c_vp]),
    'curla_agent_create': (c_vp, [C.POINTER(AgentConfig)]),
    'curla_agent_destroy': (None, [c_vp]),
    'curla_agent_arena_bytes': (c_ll, [c_vp, _i]),
    'curla_agent_bind': (_i, [c_vp, C.POINTER(c_vp)]),
    'curla_agent_num_tensors': (_i, [c_vp]),
    'curla_agent_tensor_info': (_i, [c_vp, _i, C.c_char_p, _i, C.POINTER(_i), C.POINTER(c_ll),
                                     C.POINTER(_i), C.POINTER(c_ll), C.POINTER(_i)]),
    'curla_agent_refresh_shadows': (_i, [c_vp, c_vp]),
    'curla_agent_update': (_i, [c_vp, C.POINTER(UpdateArgs), c_vp]),
    'curla_agent_last_launches': (_i, [c_vp]),
    'curla_agent_set_opt_steps': (_i, [c_vp, _i, _i, _i, _i]),
    'curla_agent_get_opt_steps': (_i, [c_vp, C.POINTER(_i)]),
    'curla_profile_enable': (_i, [_i]),
    'curla_profile_read': (_i, [C.c_char_p, _i]),
    'curla_agent_encode': (_i, [c_vp, _i, c_vp, _i, _i, c_vp, c_vp]),
    'curla_agent_actor_head': (_i, [c_vp, c_vp, _i, c_vp, C.c_ulonglong, C.c_ulonglong, _i, _i, c_vp, c_vp,
                                    c_vp, c_vp, c_vp]),
    'curla_agent_q_heads': (_i, [c_vp, _i, c_vp, c_vp, _i, c_vp, c_vp, c_vp]),
    'curla_nccl_unique_id': (_i, [c_vp]),
    'curla_agent_init_comm': (_i, [c_vp, c_vp]),
    'curla_agent_take_comm': (_i, [c_vp, c_vp]),
}

_lib = None


class CurlaError(RuntimeError):
    pass


def load():
    """Load the shared library (building is a separate, explicit step)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CurlaError('%s not found: run `python -m curla_b200.build` (there is no CPU '
                         'fallback for the CUDA hot path)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().curla_last_error()
        raise CurlaError('%s failed: %s' % (what or 'curla call', msg.decode() if msg else rc))


def ptr(t):
    """Device (or host) pointer of a torch tensor, None -> NULL."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    check(rc, name)
