"""Synthetic satellite-track observations generated on the device (vggp_generate_tracks, csrc/metrics.cuh).

The reference's `SimulationDataHour.generate_track` (src/utils/dataloaders.py:290-377) samples tracks out of a NetCDF simulation
that is not shipped; what is mirrored here is its track GEOMETRY (ascending passes x1 = o_j + t / g, x2 = t, then descending
ones, offsets o_j = j / passes, wrap-around in x1; notebook call trajectory_gradient = 2, 6_gulf_stream_experiement.ipynb:93) on
the unit square the notebooks min-max scale to, with a smooth synthetic field + noise as sea-surface height.  Every observation
is a pure function of its global index and the seed, so each rank of a data-parallel run generates its own shard.
There is no CPU path."""
import ctypes as C
from typing import List, Tuple

import torch

from .. import _lib


def generate_tracks(lo: int, hi: int, n_total: int, device, dtype=torch.float32, seed: int = 0, D: int = 2,
                    passes: int = 1024, trajectory_gradient: float = 2.0) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """Observations lo..hi-1 (acquisition order) of the synthetic track data set of `n_total` observations:
    ([x_1, .., x_D], y) as structure-of-arrays tensors of `dtype` on `device` (D = 3: acquisition time as x_3)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("generate_tracks needs a CUDA device: this package has no CPU path")
    if dtype not in (torch.float32, torch.float64):
        raise ValueError("dtype must be float32 or float64")
    n = int(hi) - int(lo)
    lib = _lib.load()
    with torch.cuda.device(device):
        xs = [torch.empty(n, dtype=dtype, device=device) for _ in range(D)]
        y = torch.empty(n, dtype=dtype, device=device)
        ptrs = (C.c_void_p * D)(*[t.data_ptr() for t in xs])
        _lib.check(lib.vggp_generate_tracks(_lib.F32 if dtype == torch.float32 else _lib.F64, int(D), int(n_total), int(lo), int(hi),
                                            int(seed), int(passes), float(trajectory_gradient), ptrs, y.data_ptr(),
                                            torch.cuda.current_stream(device).cuda_stream))
    return xs, y
