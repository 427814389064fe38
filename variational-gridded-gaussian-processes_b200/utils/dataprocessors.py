"""Min-max scaling with the reference's names and signatures (src/utils/dataprocessors.py:3-44), on CUDA tensors through
vggp_minmax / vggp_minmax_scale (csrc/metrics.cuh): one fused min+max reduction and one elementwise pass, min and max
staying on the device.  Results are bit-identical to the reference's torch expressions in the tensor's dtype.
There is no CPU path."""
import torch

from .. import _lib


def _prep(tensor: torch.Tensor):
    if tensor.device.type != "cuda":
        raise RuntimeError("min-max scaling needs a CUDA tensor: this package has no CPU path")
    if tensor.dtype not in (torch.float32, torch.float64):
        raise ValueError("tensor must be float32 or float64")
    return tensor.contiguous(), (_lib.F32 if tensor.dtype == torch.float32 else _lib.F64)


def _scalar(v, like: torch.Tensor) -> torch.Tensor:
    return v.to(like.device, like.dtype).reshape(()) if torch.is_tensor(v) else torch.tensor(v, dtype=like.dtype, device=like.device)


def min_max_scaling(tensor: torch.Tensor, min=None, max=None):
    """Returns (scaled tensor, min, max); min / max are 0-dim tensors on the tensor's device, computed when not given."""
    t, code = _prep(tensor)
    lib = _lib.load()
    with torch.cuda.device(t.device):
        st = torch.cuda.current_stream(t.device).cuda_stream
        mm = torch.empty(2, dtype=t.dtype, device=t.device)
        if min is None or max is None:
            _lib.check(lib.vggp_minmax(code, t.data_ptr(), t.numel(), mm.data_ptr(), st))
        if min is not None:
            mm[0] = _scalar(min, t)
        if max is not None:
            mm[1] = _scalar(max, t)
        out = torch.empty_like(t)
        _lib.check(lib.vggp_minmax_scale(code, t.data_ptr(), t.numel(), mm.data_ptr(), 0, out.data_ptr(), st))
    return out.view(tensor.shape), mm[0], mm[1]


def min_max_inverse(tensor: torch.Tensor, min, max):
    t, code = _prep(tensor)
    lib = _lib.load()
    with torch.cuda.device(t.device):
        st = torch.cuda.current_stream(t.device).cuda_stream
        mm = torch.stack([_scalar(min, t), _scalar(max, t)])
        out = torch.empty_like(t)
        _lib.check(lib.vggp_minmax_scale(code, t.data_ptr(), t.numel(), mm.data_ptr(), 1, out.data_ptr(), st))
    return out.view(tensor.shape)
