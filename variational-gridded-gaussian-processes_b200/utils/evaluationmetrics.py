"""Evaluation metrics with the reference's names, signatures and shape checks (src/utils/evaluationmetrics.py:6-54),
computed by one fused CUDA reduction (vggp_metrics, csrc/metrics.cuh) instead of four separate torch passes.

`true` and `pred` are 2-D CUDA tensors of the same shape and dtype (float32 or float64); each function returns a
0-dim float64 CUDA tensor.  All four metrics come from the same four sums: call `all_metrics` to get them from a single
pass over the data.  There is no CPU path."""
import torch

from .. import _lib


def _check(true: torch.Tensor, pred: torch.Tensor):
    assert len(true.shape) == 2, "true tensor must be 2D, got {}D".format(len(true.shape))
    assert len(pred.shape) == 2, "pred tensor must be 2D, got {}D".format(len(pred.shape))
    assert true.shape == pred.shape, "true and pred must have the same shape, got {} and {}".format(true.shape, pred.shape)
    if true.device.type != "cuda" or pred.device != true.device:
        raise RuntimeError("evaluation metrics need CUDA tensors on one device: this package has no CPU path")
    if true.dtype != pred.dtype or true.dtype not in (torch.float32, torch.float64):
        raise ValueError("true and pred must both be float32 or both be float64")


def metric_sums(true: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    """Device tensor [sum (t-p)^2, sum |t-p|, sum (t-t0), sum (t-t0)^2] (t0 = first target), float64."""
    _check(true, pred)
    lib = _lib.load()
    t = true.contiguous()
    p = pred.contiguous()
    out = torch.empty(4, dtype=torch.float64, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(lib.vggp_metrics(_lib.F32 if t.dtype == torch.float32 else _lib.F64, t.data_ptr(), p.data_ptr(),
                                    t.numel(), out.data_ptr(), torch.cuda.current_stream(t.device).cuda_stream))
    return out


def finish(sums: torch.Tensor, n: int) -> dict:
    """MSE, MAE, RMSE, R^2 from the four sums (0-dim float64 tensors on the sums' device)."""
    mse = sums[0] / n
    tss = sums[3] - sums[2] * sums[2] / n
    return {"mse": mse, "mae": sums[1] / n, "rmse": torch.sqrt(mse), "r2": 1 - sums[0] / tss}


def all_metrics(true: torch.Tensor, pred: torch.Tensor) -> dict:
    return finish(metric_sums(true, pred), true.numel())


def mean_squared_error(true: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    """ MSE """
    return all_metrics(true, pred)["mse"]


def mean_absolute_error(true: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    """ MAE """
    return all_metrics(true, pred)["mae"]


def root_mean_squared_error(true: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    """ RMSE """
    return all_metrics(true, pred)["rmse"]


def r_squared(true: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    """ R^2 """
    return all_metrics(true, pred)["r2"]
