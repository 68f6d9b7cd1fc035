"""ctypes loader for the C-ABI shared library (csrc/ -> libpcgrl_b200.so) and thin torch glue.

The product path has NO CPU fallback: if the library is missing or CUDA is unavailable every entry
point raises.  torch is used only to own device memory and streams; all signatures below pass raw
pointers to the ``extern "C"`` functions declared in include/pcgrl_b200.h.
"""
import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCGRL_B200_LIB") or os.path.join(_HERE, "csrc", "libpcgrl_b200.so")  # override: A/B builds

EXPORTS = ["pcgrl_abi_version", "pcgrl_last_error", "pcgrl_config_validate", "pcgrl_scratch_bytes",
           "pcgrl_reset", "pcgrl_step", "pcgrl_rollout", "pcgrl_get_stats", "pcgrl_seed", "pcgrl_step_host",
           "pcgrl_host_staging_bytes", "pcgrl_obs_image", "pcgrl_action_map", "pcgrl_rollout_host",
           "pcgrl_smb_scratch_bytes", "pcgrl_smb_get_stats", "pcgrl_reset_cpu", "pcgrl_step_cpu", "pcgrl_get_stats_cpu",
           "pcgrl_step_host_begin", "pcgrl_step_host_end", "pcgrl_render",
           "pcgrl_linear_bf16", "pcgrl_linear_last_error", "pcgrl_linear_bf16_ex", "pcgrl_im2col",
           "pcgrl_conv3x3_bf16", "pcgrl_linear_bf16_pad"]

_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    """Load libpcgrl_b200.so (built by ``__graft_entry__.build()`` / ``python -m gym_pcgrl_b200.build``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                "CUDA extension %s is missing -- build it with `python -m gym_pcgrl_b200.build` "
                "(there is no CPU fallback for the step path)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.pcgrl_abi_version.restype = C.c_int
        L.pcgrl_last_error.restype = C.c_char_p
        L.pcgrl_config_validate.restype = C.c_int
        L.pcgrl_config_validate.argtypes = [C.POINTER(_abi.PcgrlConfig)]
        L.pcgrl_scratch_bytes.restype = C.c_size_t
        L.pcgrl_scratch_bytes.argtypes = [C.POINTER(_abi.PcgrlConfig), C.c_int]
        L.pcgrl_reset.restype = C.c_int
        L.pcgrl_reset.argtypes = [C.POINTER(_abi.PcgrlConfig), C.POINTER(_abi.PcgrlBuffers), C.c_void_p, C.c_int, C.c_void_p]
        L.pcgrl_step.restype = C.c_int
        L.pcgrl_step.argtypes = [C.POINTER(_abi.PcgrlConfig), C.POINTER(_abi.PcgrlBuffers), C.c_void_p, C.c_int, C.c_void_p]
        L.pcgrl_rollout.restype = C.c_int
        L.pcgrl_rollout.argtypes = [C.POINTER(_abi.PcgrlConfig), C.POINTER(_abi.PcgrlBuffers), C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.pcgrl_get_stats.restype = C.c_int
        L.pcgrl_get_stats.argtypes = [C.POINTER(_abi.PcgrlConfig), C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.pcgrl_seed.restype = C.c_int
        L.pcgrl_seed.argtypes = [C.POINTER(_abi.PcgrlBuffers), C.c_void_p, C.c_int, C.c_void_p]
        L.pcgrl_step_host.restype = C.c_int
        L.pcgrl_step_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.pcgrl_step_host_begin.restype = C.c_int
        L.pcgrl_step_host_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.pcgrl_step_host_end.restype = C.c_int
        L.pcgrl_step_host_end.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.pcgrl_rollout_host.restype = C.c_int
        L.pcgrl_rollout_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_int, C.c_void_p]
        L.pcgrl_smb_scratch_bytes.restype = C.c_size_t
        L.pcgrl_smb_scratch_bytes.argtypes = [C.c_int, C.c_int]
        L.pcgrl_smb_get_stats.restype = C.c_int
        L.pcgrl_smb_get_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_size_t, C.c_void_p]
        L.pcgrl_obs_image.restype = C.c_int
        L.pcgrl_obs_image.argtypes = [C.POINTER(_abi.PcgrlConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.pcgrl_action_map.restype = C.c_int
        L.pcgrl_action_map.argtypes = [C.POINTER(_abi.PcgrlConfig), C.POINTER(_abi.PcgrlBuffers), C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_void_p]
        L.pcgrl_reset_cpu.restype = C.c_int
        L.pcgrl_reset_cpu.argtypes = [C.POINTER(_abi.PcgrlConfig), C.POINTER(_abi.PcgrlBuffers), C.c_void_p, C.c_int]
        L.pcgrl_step_cpu.restype = C.c_int
        L.pcgrl_step_cpu.argtypes = [C.POINTER(_abi.PcgrlConfig), C.POINTER(_abi.PcgrlBuffers), C.c_void_p, C.c_int]
        L.pcgrl_get_stats_cpu.restype = C.c_int
        L.pcgrl_get_stats_cpu.argtypes = [C.POINTER(_abi.PcgrlConfig), C.c_void_p, C.c_void_p, C.c_int]
        L.pcgrl_render.restype = C.c_int
        L.pcgrl_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p]
        L.pcgrl_linear_bf16.restype = C.c_int
        L.pcgrl_linear_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.pcgrl_linear_last_error.restype = C.c_char_p
        L.pcgrl_linear_bf16_ex.restype = C.c_int
        L.pcgrl_linear_bf16_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.pcgrl_im2col.restype = C.c_int
        L.pcgrl_im2col.argtypes = [C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p]
        L.pcgrl_conv3x3_bf16.restype = C.c_int
        L.pcgrl_conv3x3_bf16.argtypes = [C.c_void_p] * 4 + [C.c_int] * 6 + [C.c_void_p]
        L.pcgrl_linear_bf16_pad.restype = C.c_int
        L.pcgrl_linear_bf16_pad.argtypes = [C.c_void_p] * 4 + [C.c_int] * 6 + [C.c_void_p]
        L.pcgrl_host_staging_bytes.restype = C.c_size_t
        L.pcgrl_host_staging_bytes.argtypes = [C.POINTER(_abi.PcgrlConfig), C.c_int]
        if L.pcgrl_abi_version() != _abi.ABI_VERSION:
            raise NativeError("libpcgrl_b200.so ABI %d != python ABI %d" % (L.pcgrl_abi_version(), _abi.ABI_VERSION))
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().pcgrl_last_error()
        raise NativeError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else ""))


def require_cuda(device):
    import torch
    if not torch.cuda.is_available():
        raise NativeError("gym_pcgrl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise NativeError("device must be a CUDA device, got %r" % (device,))
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def validate(cfg):
    check(lib().pcgrl_config_validate(C.byref(cfg)), "pcgrl_config_validate")


def alloc_buffers(cfg, n, device):
    """Allocate the caller-owned state tensors of n envs on `device` and the struct pointing at them."""
    import torch
    h, w = cfg.height, cfg.width
    tens = {}
    for name, dtype, shape in _abi.BUFFER_SPECS:
        if name == "heatmap" and (cfg.flags & _abi.FLAG_HEAT_U16):
            dtype = "int16"   # uint16 counts (<= max_changes <= 65535 / 2 in practice) in torch's signed 16-bit storage
        tens[name] = torch.zeros((n,) + shape(h, w), dtype=getattr(torch, "int32" if dtype == "uint32" else dtype), device=device)
    tens["tile_prob"][:] = torch.tensor(list(cfg.tile_prob), dtype=torch.float64, device=device)[None, :]
    tens["status"] = torch.zeros(4, dtype=torch.int32, device=device)
    nbytes = int(lib().pcgrl_scratch_bytes(C.byref(cfg), n))
    tens["scratch"] = torch.zeros(max(nbytes, 16), dtype=torch.uint8, device=device)
    b = _abi.PcgrlBuffers()
    for name, _, _ in _abi.BUFFER_SPECS:
        setattr(b, name, tens[name].data_ptr())
    b.scratch, b.scratch_bytes, b.status = tens["scratch"].data_ptr(), nbytes, tens["status"].data_ptr()
    return tens, b


SMB_STAT_NAMES = ["dist-floor", "disjoint-tubes", "enemies", "empty", "noise", "jumps", "jumps-dist", "dist-win"]


def smb_get_stats(maps, solver_power=10000):
    """Stand-alone batched SMBProblem.get_stats (pcgrl_smb_get_stats): uint8 CUDA [N,H,W] -> int32 [N, MAX_STATS]
    (columns SMB_STAT_NAMES).  smb_prob.py:126-148."""
    import torch
    dev = require_cuda(maps.device)
    maps = maps.to(torch.uint8).contiguous()
    n, h, w = maps.shape
    out = torch.zeros((n, _abi.MAX_STATS), dtype=torch.int32, device=dev)
    nbytes = int(lib().pcgrl_smb_scratch_bytes(n, int(solver_power)))
    scratch = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib().pcgrl_smb_get_stats(maps.data_ptr(), out.data_ptr(), n, w, h, int(solver_power), scratch.data_ptr(),
                                        nbytes, stream_ptr(dev)), "pcgrl_smb_get_stats")
    return out


def get_stats_cpu(prob, maps):
    """Host twin of get_stats (pcgrl_get_stats_cpu): uint8 CPU tensor / array [N,H,W] -> int32 CPU tensor [N, MAX_STATS]."""
    import numpy as np
    import torch
    from ._config import build_config
    from .envs.reps import REPRESENTATIONS
    m = np.ascontiguousarray(np.asarray(maps), dtype=np.uint8)
    n, h, w = m.shape
    if (h, w) != (prob._height, prob._width):
        raise ValueError("map shape %s does not match the problem's (height, width) = %s" % ((h, w), (prob._height, prob._width)))
    cfg = build_config(prob, REPRESENTATIONS["wide"](), 1, 1, auto_reset=False)
    out = np.zeros((n, _abi.MAX_STATS), dtype=np.int32)
    check(lib().pcgrl_get_stats_cpu(C.byref(cfg), m.ctypes.data, out.ctypes.data, n), "pcgrl_get_stats_cpu")
    return torch.from_numpy(out)


def get_stats(prob, maps):
    """Stand-alone batched Problem.get_stats (pcgrl_get_stats): uint8 CUDA [N,H,W] -> int32 [N, MAX_STATS].
    A CPU tensor goes to the host twin."""
    import torch
    from ._config import build_config
    from .envs.reps import REPRESENTATIONS
    if isinstance(maps, torch.Tensor) and maps.device.type == "cpu":
        return get_stats_cpu(prob, maps)
    dev = require_cuda(maps.device)
    n, h, w = maps.shape
    if (h, w) != (prob._height, prob._width):
        raise ValueError("map shape %s does not match the problem's (height, width) = %s" % ((h, w), (prob._height, prob._width)))
    cfg = build_config(prob, REPRESENTATIONS["wide"](), 1, 1, auto_reset=False)
    validate(cfg)
    maps = maps.to(torch.uint8).contiguous()
    out = torch.zeros((n, _abi.MAX_STATS), dtype=torch.int32, device=dev)
    nbytes = int(lib().pcgrl_scratch_bytes(C.byref(cfg), n))
    scratch = torch.zeros(max(nbytes, 16), dtype=torch.uint8, device=dev)
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib().pcgrl_get_stats(C.byref(cfg), maps.data_ptr(), out.data_ptr(), n, scratch.data_ptr(), nbytes,
                                    status.data_ptr(), stream_ptr(dev)), "pcgrl_get_stats")
    if int(status[0].item()) != 0:
        raise NativeError("pcgrl_get_stats hit a device capacity limit (status=%s)" % status.tolist())
    return out


def render(prob, maps, pos=None):
    """pcgrl_render: batched PcgrlEnv.render(mode='rgb_array') -> uint8 CUDA tensor [N, Hpx, Wpx, 3]."""
    import torch
    dev = require_cuda(maps.device)
    maps = maps.to(torch.uint8).contiguous()
    n, h, w = maps.shape
    atlas = torch.as_tensor(prob.get_graphics()).to(device=dev, dtype=torch.uint8).contiguous()
    tiles = prob.get_tile_types()
    ts, (bw, bh) = int(prob._tile_size), prob._border_size
    out = torch.empty((n, (h + 2 * bh) * ts, (w + 2 * bw) * ts, 3), dtype=torch.uint8, device=dev)
    p = None if pos is None else pos.to(torch.uint8).contiguous()
    with torch.cuda.device(dev):
        check(lib().pcgrl_render(maps.data_ptr(), None if p is None else p.data_ptr(), atlas.data_ptr(), out.data_ptr(), n, h, w,
                                 len(tiles), int(bw), int(bh), tiles.index(prob._border_tile), ts, stream_ptr(dev)), "pcgrl_render")
    return out


def linear_bf16(x, weight, bias=None, relu=True):
    """pcgrl_linear_bf16 (csrc/pcgrl_linear.cu, tcgen05 + TMEM + TMA): relu(x @ weight.T + bias) -> float32 [M, N].
    x [M, K] and weight [N, K] are converted to contiguous bf16 if they are not already."""
    import torch
    dev = require_cuda(x.device)
    xb = x.to(torch.bfloat16).contiguous()
    wb = weight.to(device=dev, dtype=torch.bfloat16).contiguous()
    m, k = xb.shape
    n = wb.shape[0]
    b = None if bias is None else bias.to(device=dev, dtype=torch.float32).contiguous()
    y = torch.empty((m, n), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib().pcgrl_linear_bf16(xb.data_ptr(), wb.data_ptr(), None if b is None else b.data_ptr(), y.data_ptr(), m, n, k,
                                     1 if relu else 0, stream_ptr(dev))
    if rc:
        raise NativeError("pcgrl_linear_bf16 failed (rc=%d): %s" % (rc, lib().pcgrl_linear_last_error().decode()))
    return y
