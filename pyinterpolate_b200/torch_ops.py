"""``torch.ops.pyinterpolate_b200.*`` — the device-resident entry points of libpyinterp_b200.so registered as PyTorch custom
operators (torch.library), for callers whose points already live on the GPU as tensors:

* ``torch.ops.pyinterpolate_b200.variogram_sums(xyz, edges, lower, W, shard_index, shard_count)``
  -> ``(sums[3, (D,) L] float64, count[(D,) L] int64)``: the pair pass behind ``ExperimentalVariogram`` /
  ``calculate_semivariance`` / ``calculate_covariance`` (semivariogram/experimental/functions/general.py:235-303,
  functions/directional.py:395-465); ``sums`` = sum (z_i - z_j)^2, sum z_i z_j, sum (z_i + z_j) per lag over unordered pairs.
* ``torch.ops.pyinterpolate_b200.krige(known, unknown, params, model, mode, process_mean, no_neighbors)``
  -> ``(out[(2,) S, M, 4] float64 rows [zhat, sigma^2, x, y], status[(2,) S, M] uint8)``: neighbour search, assembly and solve
  behind ``ordinary_kriging`` / ``simple_kriging`` (kriging/point/ordinary.py:134-221, simple.py:129-223).

Both run on the tensor's device and torch's current stream through the C ABI (``pib_variogram_device``, ``pib_krige_device``);
``edges`` / ``lower`` / ``W`` / ``params`` are small host-side float64 tensors (lag thresholds, whitening matrices, model
parameters).  No autograd (the reference's functions have none), no CPU implementation: a CPU tensor raises.
``register()`` defines the operators (idempotent); nothing is registered on a plain ``import pyinterpolate_b200``."""
import torch

from . import _lib

_registered = False


def register():
    """Define the operators once (idempotent)."""
    global _registered
    if _registered:
        return
    _registered = True

    @torch.library.custom_op("pyinterpolate_b200::variogram_sums", mutates_args=(), device_types="cuda")
    def variogram_sums(xyz: torch.Tensor, edges: torch.Tensor, lower: torch.Tensor, W: torch.Tensor, shard_index: int,
                       shard_count: int) -> tuple[torch.Tensor, torch.Tensor]:
        r = _lib.variogram_device(xyz.contiguous(), edges.cpu().numpy(), lower=lower.cpu().numpy() if lower.numel() else None,
                                  W=W.cpu().numpy() if W.numel() else None, shard=(shard_index, shard_count))
        return torch.stack([r["sum_sq"], r["sum_prod"], r["sum_val"]]), r["count"]

    @variogram_sums.register_fake
    def _(xyz, edges, lower, W, shard_index, shard_count):
        shape = (edges.shape[0],) if W.numel() == 0 else (W.reshape(-1, 4).shape[0], edges.shape[0])
        return xyz.new_empty((3,) + shape), xyz.new_empty(shape, dtype=torch.int64)

    @torch.library.custom_op("pyinterpolate_b200::krige", mutates_args=(), device_types="cuda")
    def krige(known: torch.Tensor, unknown: torch.Tensor, params: torch.Tensor, model: int, mode: int, process_mean: float,
              no_neighbors: int) -> tuple[torch.Tensor, torch.Tensor]:
        out, _, status = _lib.krige_device(known.contiguous(), unknown.contiguous(), model, params.cpu().numpy(), mode,
                                           no_neighbors, process_mean=process_mean)
        return out, status

    @krige.register_fake
    def _(known, unknown, params, model, mode, process_mean, no_neighbors):
        lead = (2,) if mode == _lib.KRIGE_BOTH else ()
        shape = lead + (params.reshape(-1, 3).shape[0], unknown.shape[0])
        return known.new_empty(shape + (4,)), known.new_empty(shape, dtype=torch.uint8)


def empty_f64():
    """The 'not given' value of the optional tensor arguments (``lower``, ``W``)."""
    return torch.zeros(0, dtype=torch.float64)


def model_id(name):
    return _lib.MODEL_IDS[name]


__all__ = ["register", "empty_f64", "model_id"]
