"""numpy restatement of the counter-based noise stream used by the CUDA sampler kernels
(TEST INFRASTRUCTURE).

The reference draws noise from torch's global generator (`torch.randn`, `randn_like`,
score_sampling.py:94,125,168,204,227; score_unet.py:957-959), whose stream cannot be
reproduced inside a fused kernel.  Parity tests therefore *inject* noise: both the oracle
samplers and the CUDA kernels consume this stream.

Stream definition (the contract `include/sbgm_b200.h` documents for `sbgm_philox_normal`):
  Philox4x32-10, key = (seed & 0xffffffff, seed >> 32),
  counter = (q & 0xffffffff, q >> 32, draw, 0) where q = element_index // 4;
  the four 32-bit outputs r0..r3 give four normals for elements 4q..4q+3:
    u = ((r >> 8) + 0.5) * 2^-24;   rad = sqrt(-2 ln u_a);   ang = 2 pi u_b
    z[4q+0] = rad(r0) cos(ang(r1)), z[4q+1] = rad(r0) sin(ang(r1)),
    z[4q+2] = rad(r2) cos(ang(r3)), z[4q+3] = rad(r2) sin(ang(r3)).
  Uniform draws use u(r0..r3) directly.
`element_index` is the index into the *global* [members, C, H, W] tensor, so a sharded
ensemble reproduces the single-GPU stream exactly.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr: [n,4] uint32, key: [2] uint32 -> [n,4] uint32."""
    c0, c1, c2, c3 = (ctr[:, i].astype(np.uint64) for i in range(4))
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK
        n2 = (hi0 ^ c3 ^ np.uint64(k1)) & MASK
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        with np.errstate(over="ignore"):
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def _raw(n_elems: int, seed: int, draw: int, first_elem: int = 0) -> np.ndarray:
    assert first_elem % 4 == 0
    nq = (n_elems + 3) // 4
    q = np.arange(nq, dtype=np.uint64) + np.uint64(first_elem // 4)
    ctr = np.zeros((nq, 4), dtype=np.uint32)
    ctr[:, 0] = (q & MASK).astype(np.uint32)
    ctr[:, 1] = (q >> np.uint64(32)).astype(np.uint32)
    ctr[:, 2] = np.uint32(draw)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return philox4x32_10(ctr, key)


def _u01(r: np.ndarray) -> np.ndarray:
    return ((r >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)


def uniform(n_elems: int, seed: int, draw: int, first_elem: int = 0) -> np.ndarray:
    return _u01(_raw(n_elems, seed, draw, first_elem)).reshape(-1)[:n_elems]


def normal(n_elems: int, seed: int, draw: int, first_elem: int = 0) -> np.ndarray:
    """float32 standard normals for elements first_elem .. first_elem + n_elems."""
    u = _u01(_raw(n_elems, seed, draw, first_elem))
    out = np.empty_like(u)
    for a, b in ((0, 1), (2, 3)):
        rad = np.sqrt(np.float32(-2.0) * np.log(u[:, a])).astype(np.float32)
        ang = (np.float32(6.2831855) * u[:, b]).astype(np.float32)
        out[:, a] = rad * np.cos(ang)
        out[:, b] = rad * np.sin(ang)
    return out.reshape(-1)[:n_elems].astype(np.float32)


# draw ids -- shared with sbgm_danra_b200/csrc (see include/sbgm_b200.h)
DRAW_INIT = 0


def draw_em(step: int) -> int:
    return 1 + step


def draw_pc_corrector(step: int) -> int:
    return 1 + 2 * step


def draw_pc_predictor(step: int) -> int:
    return 2 + 2 * step


DRAW_DSM_T = 0
DRAW_DSM_Z = 1
