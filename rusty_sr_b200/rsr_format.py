"""The .rsr parameter-file codec: bytevec 0.2 `<Vec<f32>>::decode::<u32>` / `.encode::<u32>()`
as used at reference src/main.rs:138,146,149,152 (decode) and :213 (encode).

    u32 LE  N | N x u32 LE element byte sizes (all 4) | N x f32 LE        
"""
from __future__ import annotations

import struct

import numpy as np


def decode_rsr(data: bytes) -> np.ndarray:
    """Returns the flat f32 parameter vector.  Raises ValueError("ByteVec conversion failed")
    on malformed input — the message of the reference's `.expect(..)` (main.rs:138)."""
    if len(data) < 4:
        raise ValueError("ByteVec conversion failed")
    (n,) = struct.unpack_from("<I", data, 0)
    if len(data) != 4 + 8 * n:
        raise ValueError("ByteVec conversion failed")
    sizes = np.frombuffer(data, dtype="<u4", count=n, offset=4)
    if n and not (sizes == 4).all():
        raise ValueError("ByteVec conversion failed")
    return np.frombuffer(data, dtype="<f4", count=n, offset=4 + 4 * n).astype(np.float32)


def encode_rsr(params) -> bytes:
    """Inverse of decode_rsr (the format training checkpoints are written in, main.rs:213)."""
    p = np.ascontiguousarray(params, dtype="<f4").ravel()
    n = p.size
    return struct.pack("<I", n) + np.full(n, 4, dtype="<u4").tobytes() + p.tobytes()
