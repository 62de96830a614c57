"""ctypes binding of ``libforagax_b200.so`` (the C ABI declared in ``include/foragax_b200.h``).

There is no CPU fallback: if the shared library is missing, importing this module raises.
PyTorch is used only as plumbing -- it owns device memory (``tensor.data_ptr()``) and the
CUDA stream the kernels are enqueued on.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libforagax_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`) "
        "-- foragax_b200 has no CPU fallback")

lib = C.CDLL(LIB_PATH)

FGX_U8, FGX_I32, FGX_F32, FGX_U32 = 0, 1, 2, 3
OP_EQ, OP_NE, OP_LT, OP_LE, OP_GT, OP_GE, OP_NONZERO = range(7)
MAX_LEAVES = 32

vp, i64, i32, u32, f32, sz = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_float, C.c_size_t


class PredTerm(C.Structure):
    _fields_ = [("data", vp), ("dtype", i32), ("op", i32), ("imm_bits", u32), ("stride", i32)]


class Predicate(C.Structure):
    _fields_ = [("term", PredTerm * 4), ("n_terms", i32), ("lut", u32)]


class Leaf(C.Structure):
    _fields_ = [("src", vp), ("dst", vp), ("row_bytes", i64)]


MAX_PEERS = 8
MAX_PEER_LEAVES = 16


class PeerLeaf(C.Structure):
    _fields_ = [("src", vp), ("dst", vp * MAX_PEERS), ("row_bytes", i64)]


class Walls(C.Structure):
    _fields_ = [("x1", vp), ("y1", vp), ("x2", vp), ("y2", vp), ("n", i64)]


def _decl(name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_decl("fgx_last_error", C.c_char_p)
_decl("fgx_version", C.c_int)
_decl("fgx_device_info", C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int))
_decl("fgx_threefry_split", C.c_int, vp, i64, i32, vp, vp)
_decl("fgx_threefry_random_bits", C.c_int, vp, i64, i32, vp, vp)
_decl("fgx_threefry_uniform", C.c_int, vp, i64, i32, f32, f32, vp, vp)
_decl("fgx_threefry_randint", C.c_int, vp, i64, i32, i32, i32, vp, vp)
_decl("fgx_threefry_normal", C.c_int, vp, i64, i32, vp, vp)
_decl("fgx_select_scratch_bytes", sz, i64)
_decl("fgx_select", C.c_int, C.POINTER(Predicate), i64, vp, vp, vp, sz, vp)
_decl("fgx_argsort_scratch_bytes", sz, i64)
_decl("fgx_argsort", C.c_int, vp, i32, i64, i32, vp, vp, sz, vp)
_decl("fgx_sort_rows", C.c_int, vp, i32, i64, i32, vp, C.POINTER(Leaf), i32, vp, sz, vp)
_decl("fgx_gather_rows", C.c_int, C.POINTER(Leaf), i32, vp, i64, vp)
_decl("fgx_gather_rows_from", C.c_int, C.POINTER(Leaf), i32, vp, i64, i64, vp)
_decl("fgx_pack_rows_peer", C.c_int, C.POINTER(PeerLeaf), i32, vp, vp, i64, vp, C.POINTER(i64), i32, i32, vp)


class PeerSync(C.Structure):
    _fields_ = [("mailbox", vp * MAX_PEERS), ("epoch", vp), ("error", vp), ("rank", i32), ("world", i32)]


_decl("fgx_peer_mailbox_bytes", sz)
_decl("fgx_peer_exchange", C.c_int, C.POINTER(PeerSync), vp, vp, vp, vp)
_decl("fgx_peer_gather_positions", C.c_int, C.POINTER(PeerSync), C.POINTER(vp), i64, vp, vp, vp, vp, i64, i64, i32,
      C.POINTER(i64), vp, vp)
_decl("fgx_peer_add_range", C.c_int, C.POINTER(PeerSync), vp, vp, i64, i64, i64, vp, vp)
_decl("fgx_key_chain_advance", C.c_int, i32, vp, vp, vp, vp)
_decl("fgx_sort_key_images", C.c_int, vp, i32, i64, i32, vp, vp, vp)
_decl("fgx_sort_rows_peer", C.c_int, C.POINTER(PeerLeaf), i32, vp, i64, vp, i64, C.POINTER(i64), i32, i32, vp, vp)
_decl("fgx_sum_f32_scratch_bytes", sz)
_decl("fgx_sum_f32", C.c_int, vp, i64, vp, vp, sz, vp)
_decl("fgx_life_expectancy", C.c_int, vp, vp, vp, vp, vp, vp)
_decl("fgx_record_frame", C.c_int, C.POINTER(Leaf), i32, i64, i32, vp, vp)
_decl("fgx_count_active", C.c_int, vp, i64, vp, vp)
_decl("fgx_count_active_clip", C.c_int, vp, i64, vp, vp, vp)
_decl("fgx_create_base", C.c_int, C.POINTER(u32), i64, i64, i32, vp, vp, vp, vp, vp)
_decl("fgx_create_base_shard", C.c_int, C.POINTER(u32), i64, i64, i64, i64, i64, i32, vp, vp, vp, vp, vp)
_decl("fgx_raycast_walls", C.c_int, vp, vp, vp, vp, i64, i32, f32, f32, vp, vp, vp, vp, i64, vp, vp, vp)
_decl("fgx_raycast_walls_grid_scratch_bytes", sz, i64, f32, f32, f32, f32, f32)
_decl("fgx_raycast_walls_grid", C.c_int, vp, vp, vp, vp, i64, i32, f32, f32, vp, vp, vp, vp, i64, f32, f32, f32, f32,
      vp, vp, vp, sz, i32, vp)
_decl("fgx_raycast_walls_grid_listed", C.c_int, vp, vp, vp, i64, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, vp, sz, i32,
      vp, vp)
_decl("fgx_raycast_agents_scratch_bytes", sz, i64, i64, f32, f32, f32, f32, f32, f32)
_decl("fgx_raycast_agents", C.c_int, vp, vp, vp, vp, i64, i32, f32, f32, vp, vp, vp, vp, i64, i64, f32, f32, f32, f32,
      f32, i32, vp, vp, vp, sz, vp)
_decl("fgx_ray_generator", C.c_int, vp, i64, i32, f32, vp, vp, vp)
_decl("fgx_ray_wall_sensor", C.c_int, vp, vp, vp, vp, i64, f32, vp, vp, vp, vp, i64, vp, vp)
_decl("fgx_ray_agent_sensor", C.c_int, vp, vp, vp, vp, i64, f32, vp, vp, vp, vp, i64, vp, vp)
_decl("fgx_check_collision", C.c_int, vp, vp, vp, vp, i64, vp, vp, vp, vp, i64, vp, vp)
_decl("fgx_dice_create", C.c_int, i64, i32, vp, vp, vp, vp, vp, vp)
_decl("fgx_dice_step", C.c_int, i64, i32, vp, vp, vp, vp, vp)
_decl("fgx_dice_remove", C.c_int, i64, vp, vp, vp, vp, vp)
_decl("fgx_dice_add", C.c_int, i64, i32, vp, vp, vp, vp, vp, vp)


class ForagerParams(C.Structure):
    _fields_ = [("n_neurons", i32), ("n_obs", i32), ("n_act", i32), ("dt", f32), ("policy_dt", f32),
                ("action_scale", f32), ("x_min", f32), ("x_max", f32), ("y_min", f32), ("y_max", f32),
                ("repr_thresh", f32), ("die_thresh", f32), ("energy_min", f32), ("energy_max", f32),
                ("energy_cost_dt", f32), ("clip_min", i32)]


FORAGER_STATE_FIELDS = ("X", "X_vel", "X_acc", "Y", "Y_vel", "Y_acc", "ang", "energy", "rad",
                        "time_above_rep_thresh", "time_below_die_thresh")


class ForagerState(C.Structure):
    _fields_ = [(k, vp) for k in FORAGER_STATE_FIELDS]


class WcPolicy(C.Structure):
    _fields_ = [(k, vp) for k in ("J", "tau", "E", "B", "D", "Z")]


class ResourceParams(C.Structure):
    _fields_ = [("dt", f32), ("x_min", f32), ("x_max", f32), ("blossom_thresh", f32)]


class ResourceState(C.Structure):
    _fields_ = [(k, vp) for k in ("X", "Y", "energy", "rad", "time_above_blossom", "blossom", "key")]


_P = C.POINTER
_decl("fgx_forager_step", C.c_int, _P(ForagerParams), i64, vp, vp, vp, _P(ForagerState), _P(WcPolicy), vp, vp)
_decl("fgx_forager_sense_step", C.c_int, _P(ForagerParams), i64, vp, vp, _P(ForagerState), _P(WcPolicy), vp, i32, f32,
      f32, vp, vp, vp, vp, i64, f32, f32, f32, f32, vp, sz, i32, vp, vp, vp)
_decl("fgx_wilson_cowan_step", C.c_int, i64, i32, i32, i32, f32, f32, vp, vp, _P(WcPolicy), vp, vp)
_decl("fgx_resource_step", C.c_int, _P(ResourceParams), i64, vp, vp, _P(ResourceState), vp, vp, vp, vp)
_decl("fgx_forager_resource_interaction", C.c_int, i64, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp)
_decl("fgx_sensor_model", C.c_int, i64, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp,
      i32, f32, f32, vp, vp)
_decl("fgx_forager_create", C.c_int, _P(ForagerParams), i64, vp, vp, _P(ForagerState), vp, vp, vp, vp)
_decl("fgx_forager_remove", C.c_int, i64, i32, vp, vp, _P(ForagerState), vp, vp, vp, vp)
_decl("fgx_forager_add", C.c_int, i64, i32, i32, i32, vp, vp, vp, vp, vp, f32, i32, _P(ForagerState), _P(WcPolicy),
      vp, vp, vp, vp)
_decl("fgx_forager_set", C.c_int, i64, vp, vp, _P(ForagerState), vp)
_decl("fgx_resource_create", C.c_int, i64, f32, f32, f32, f32, vp, vp, _P(ResourceState), vp, vp, vp, vp)
_decl("fgx_resource_remove", C.c_int, i64, vp, vp, _P(ResourceState), vp, vp, vp)
_decl("fgx_resource_add", C.c_int, i64, f32, f32, f32, f32, vp, vp, vp, vp, vp, _P(ResourceState), vp, vp, vp, vp, vp,
      vp)
_decl("fgx_resource_set", C.c_int, i64, vp, vp, _P(ResourceState), vp)


class AnimalState(C.Structure):
    _fields_ = [(k, vp) for k in ("X_pos", "Y_pos", "energy", "reproduce", "key", "policy_key")]


class AnimalParams(C.Structure):
    _fields_ = [("x_min", i32), ("x_max", i32), ("y_min", i32), ("y_max", i32), ("torous", i32),
                ("delta_energy", i32), ("reproduction_rate", f32)]


_decl("fgx_animal_create", C.c_int, i64, i32, i32, i32, vp, vp, _P(AnimalState), vp, vp)
_decl("fgx_animal_step", C.c_int, _P(AnimalParams), i64, vp, vp, _P(AnimalState), vp, vp)
_decl("fgx_animal_remove", C.c_int, i64, vp, vp, _P(AnimalState), vp, vp, vp)
_decl("fgx_animal_add", C.c_int, i64, vp, vp, vp, _P(AnimalState), vp, vp, vp)
_decl("fgx_animal_set", C.c_int, i64, vp, vp, _P(AnimalState), vp)
_decl("fgx_grass_create", C.c_int, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp)
_decl("fgx_grass_step", C.c_int, i64, i32, vp, vp, vp, vp)
_decl("fgx_wolf_sheep_interaction_scratch_bytes", sz, i32, i32)
_decl("fgx_wolf_sheep_interaction", C.c_int, i32, i32, i64, vp, vp, i64, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp,
      sz, vp)


class FgxError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise FgxError(lib.fgx_last_error().decode() + f" (code {rc})")


def require_cuda():
    if not torch.cuda.is_available():
        raise FgxError("foragax_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise FgxError("expected a CUDA tensor")
    if not t.is_contiguous():
        raise FgxError("expected a contiguous tensor")
    return t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cuda_ok = None
_devices = {}


def stream():
    """Raw handle of torch's current stream on the current device (every launcher enqueues on it).  Called
    once per native op: the direct binding skips the ~4 us of torch.cuda.current_stream()'s Python."""
    if _raw_stream is not None:
        return _raw_stream(torch._C._cuda_getDevice())
    return torch.cuda.current_stream().cuda_stream


def device():
    global _cuda_ok
    if _cuda_ok is None:
        require_cuda()
        _cuda_ok = True
    idx = torch._C._cuda_getDevice()
    dev = _devices.get(idx)
    if dev is None:
        dev = _devices[idx] = torch.device("cuda", idx)
    return dev


def dtype_code(t):
    if t.dtype == torch.float32:
        return FGX_F32
    if t.dtype == torch.int32:
        return FGX_I32
    if t.dtype in (torch.bool, torch.uint8):
        return FGX_U8
    if t.dtype == torch.uint32:
        return FGX_U32
    raise FgxError(f"unsupported leaf dtype {t.dtype}")


def scratch(nbytes):
    """Scratch buffer from torch's caching allocator (512-byte aligned)."""
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device())


def device_i32(v):
    """A traced count: int32 device scalar from a Python int or a device tensor."""
    if isinstance(v, torch.Tensor):
        if v.dtype != torch.int32 or not v.is_cuda:
            v = v.to(device=device(), dtype=torch.int32)
        return v.reshape(-1)[:1].contiguous()
    return torch.tensor([int(v)], dtype=torch.int32, device=device())
