"""Device-to-device copy helper for results left in engine-owned device memory
(b200aln_batch_device): lets a caller wrap them in its own tensors without
going through the host.  Uses the CUDA runtime binding of cuda-python."""
from __future__ import annotations


def d2d(dst_ptr: int, src_ptr: int, nbytes: int) -> None:
    from cuda.bindings import runtime as rt
    if nbytes <= 0:
        return
    (err,) = rt.cudaMemcpy(dst_ptr, src_ptr, nbytes, rt.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
    if int(err) != 0:
        raise RuntimeError(f"cudaMemcpy D2D failed: {err}")
