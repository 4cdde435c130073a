"""Philox4x32-10 + ATen ``normal_`` element mapping, restated in NumPy (TEST INFRASTRUCTURE ONLY).

The reference draws ``eps`` with ``logvar.data.new(size).normal_()``
(models/networks.py:230), i.e. from the default generator of logvar's device.
On CUDA that is ATen's ``distribution_elementwise_grid_stride_kernel``
(ATen/native/cuda/DistributionTemplates.h:65-91 in the installed torch 2.11
headers) with ``curand_normal4`` on a ``curandStatePhilox4_32_10_t``
(CUDA 12.9 ``curand_philox4x32_x.h``, ``curand_normal.h:70-87``):

  grid   = min(ceil(n/256), SMs * (2048/256)),  block = 256        (:50-62)
  thread = blockIdx*256 + threadIdx; curand_init(seed, subsequence=thread, offset)
  loop it = 0..: rand4 = box_muller4(philox(counter)), element li = thread + (4*it+ii)*grid*256 takes rand4[ii]
  generator offset afterwards += ((n-1)//(256*grid*4) + 1) * 4

The integer part (Philox counters -> 4x uint32) is reproduced bit-exactly
here.  The float part uses device ``logf`` / ``__sincosf`` on the GPU; the
NumPy version below evaluates the same formula in float32 with libm, so it
agrees to ~1e-6 but is not the bit-exact authority -- on the GPU box the
bit-exact check is made against ``torch.empty(n, device='cuda').normal_()``.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: uint32 [...,4], key: uint32 [...,2] -> uint32 [...,4] (10 rounds, curand_philox4x32_x.h)."""
    c = [np.asarray(ctr[..., i], np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], np.uint64)
    k1 = np.asarray(key[..., 1], np.uint64)
    for r in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [(hi1 ^ c[1] ^ k0) & MASK, lo1, (hi0 ^ c[3] ^ k1) & MASK, lo0]
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def aten_normal_policy(n, num_sms=148, max_threads_per_sm=2048):
    """(grid, counter_offset_increment) of calc_execution_policy (DistributionTemplates.h:50-62), unroll 4."""
    block = 256
    grid = min((n + block - 1) // block, num_sms * (max_threads_per_sm // block))
    inc = ((n - 1) // (block * grid * 4) + 1) * 4
    return grid, inc


def aten_normal_raw(n, seed, offset, num_sms=148):
    """uint32 pairs (x, y) that feed Box-Muller for each of the n output elements, plus which branch.

    Returns (u32 [n,2], use_cos [n] bool): element li uses box_muller(x,y).sin for
    components 0 and 2 of the float4 and .cos for 1 and 3.
    """
    grid, _ = aten_normal_policy(n, num_sms)
    nthreads = grid * 256
    li = np.arange(n, dtype=np.int64)
    thread = li % nthreads
    slot = li // nthreads            # = 4*it + ii
    it = slot // 4
    ii = slot % 4
    # curand_init(seed, subsequence, offset): counter = (offset/4 + it [64-bit in c0,c1], subsequence [64-bit in c2,c3])
    cnt = (np.uint64(offset // 4) + it.astype(np.uint64))
    ctr = np.stack([(cnt & MASK).astype(np.uint32), (cnt >> np.uint64(32)).astype(np.uint32),
                    (thread.astype(np.uint64) & MASK).astype(np.uint32),
                    (thread.astype(np.uint64) >> np.uint64(32)).astype(np.uint32)], axis=-1)
    key = np.empty((n, 2), np.uint32)
    key[:, 0] = np.uint32(seed & 0xFFFFFFFF)
    key[:, 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    out = philox4x32_10(ctr, key)
    pair = np.where((ii >= 2)[:, None], out[:, 2:4], out[:, 0:2])
    return pair, (ii % 2 == 1)


def box_muller_f32(pair, use_cos):
    """curand_normal.h:70-87 in float32 (libm instead of device intrinsics)."""
    x = pair[:, 0].astype(np.float32)
    y = pair[:, 1].astype(np.float32)
    inv = np.float32(2.3283064e-10)
    inv2pi = np.float32(2.3283064e-10) * np.float32(6.2831855)
    # x*c + c/2 is contracted to one fma by nvcc; emulate the single rounding through float64
    u = (x.astype(np.float64) * np.float64(inv) + np.float64(inv / np.float32(2))).astype(np.float32)
    v = (y.astype(np.float64) * np.float64(inv2pi) + np.float64(inv2pi / np.float32(2))).astype(np.float32)
    s = np.sqrt(np.float32(-2.0) * np.log(u, dtype=np.float32), dtype=np.float32)
    r = np.where(use_cos, np.cos(v, dtype=np.float32), np.sin(v, dtype=np.float32)).astype(np.float32)
    return (r * s).astype(np.float32)


def aten_normal(n, seed, offset, num_sms=148):
    pair, use_cos = aten_normal_raw(n, seed, offset, num_sms)
    return box_muller_f32(pair, use_cos)


# Known-answer test from the Random123 distribution (kat_vectors, philox4x32-10):
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]
