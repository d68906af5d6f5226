"""Synthetic SIFT-like descriptor sets for tests and benchmarks.

The reference pipeline obtains descriptors from its ``SiftExtraction`` op
(``integration/op_cpp/extraction_op.cc:107-118``): N x 128 uint8, each row
L2-normalised to length ~512.  Uniform-random bytes are a degenerate stand-in
(every dot product exceeds 512^2, ``acos`` saturates and the ratio test rejects
every row -- SURVEY.md 8d), so this generator produces descriptors with the
same gross statistics as real SIFT:

* components ~ Gamma(0.5, 1) (sparse, non-negative), row L2-normalised to 512,
  rounded, clamped to 255;
* a sliding pool of "track" vectors shared between neighbouring images (plus
  per-image noise), so that image pairs inside the matching window have true
  correspondences while everything else looks like clutter.

Everything is a pure function of (seed, image_id, n), so any rank of a
multi-GPU job can generate exactly its own images.
"""
from __future__ import annotations

import numpy as np

DIM = 128
_TRACK_CHUNK = 1024


def _rng(*key: int) -> np.random.Generator:
    return np.random.Generator(np.random.MT19937(np.random.SeedSequence(list(key))))


def _track_vectors(first: int, count: int, seed: int) -> np.ndarray:
    """float32 [count, 128] base vectors of tracks first .. first+count-1 (deterministic per track)."""
    out = np.empty((count, DIM), dtype=np.float32)
    pos = 0
    t = first
    while pos < count:
        chunk = t // _TRACK_CHUNK
        lo = t - chunk * _TRACK_CHUNK
        take = min(_TRACK_CHUNK - lo, count - pos)
        block = _rng(seed, 0x7261636B, chunk).gamma(0.5, 1.0, size=(_TRACK_CHUNK, DIM)).astype(np.float32)
        out[pos:pos + take] = block[lo:lo + take]
        pos += take
        t += take
    return out


def quantize(v: np.ndarray) -> np.ndarray:
    """L2-normalise rows to 512, round, clamp to [0, 255] (the SIFT uint8 convention)."""
    v = np.maximum(v, 0.0)
    nrm = np.linalg.norm(v, axis=1, keepdims=True)
    nrm[nrm == 0] = 1.0
    return np.clip(np.rint(v * (512.0 / nrm)), 0, 255).astype(np.uint8)


def make_image(image_id: int, n: int, *, seed: int = 1234, shared_frac: float = 0.4,
               track_step: int = 256, noise: float = 0.12) -> np.ndarray:
    """uint8 [n, 128] descriptors of synthetic image ``image_id``.

    ``shared_frac`` of the rows follow tracks ``image_id*track_step + [0, n_shared)``, so images
    k and k+d share ``n_shared - d*track_step`` tracks (when positive); the rest is clutter.
    Rows are randomly permuted so matching indices differ between images.
    """
    n = int(n)
    if n == 0:
        return np.empty((0, DIM), dtype=np.uint8)
    rng = _rng(seed, 0x696D6167, int(image_id))
    n_sh = int(n * shared_frac)
    parts = []
    if n_sh:
        base = _track_vectors(int(image_id) * track_step, n_sh, seed)
        parts.append(base + noise * rng.standard_normal(size=base.shape).astype(np.float32))
    if n - n_sh:
        parts.append(rng.gamma(0.5, 1.0, size=(n - n_sh, DIM)).astype(np.float32))
    v = np.concatenate(parts, axis=0)
    return np.ascontiguousarray(quantize(v)[rng.permutation(n)])


def make_images(num_images: int, n, *, first_id: int = 0, seed: int = 1234, **kw):
    """List of descriptor arrays for image ids first_id .. first_id+num_images-1.
    ``n`` is an int or a per-image sequence."""
    ns = [int(n)] * num_images if np.isscalar(n) else [int(x) for x in n]
    return [make_image(first_id + k, ns[k], seed=seed, **kw) for k in range(num_images)]


def ragged_sizes(num_images: int, lo: int = 1024, hi: int = 16384, seed: int = 1234) -> np.ndarray:
    """Log-uniform keypoint counts in [lo, hi] (BASELINE.json configs[3])."""
    u = _rng(seed, 0x72616767).random(num_images)
    return np.exp(np.log(lo) + u * (np.log(hi) - np.log(lo))).astype(np.int64)


# ------------------------------------------------------------------------------------------------------------
# The same model generated ON the GPU (torch): thousands of full-size images in a second instead of minutes, for
# the large benchmark configurations (1000 x 8192, 2000 ragged, 200 x 16384) whose descriptors are resident in HBM
# when their timed region starts.  Different random streams than make_image (so different bytes), same statistics;
# a pure function of (seed, image_id, n) on a given GPU type.
# ------------------------------------------------------------------------------------------------------------
_TORCH_TRACKS = {}


def _track_chunk_torch(chunk: int, seed: int, device):
    import torch
    key = (str(device), seed, chunk)
    t = _TORCH_TRACKS.get(key)
    if t is None:
        if len(_TORCH_TRACKS) > 64:
            _TORCH_TRACKS.clear()
        g = torch.Generator(device=device)
        g.manual_seed((seed * 1000003 + chunk) * 2 + 1)
        t = torch._standard_gamma(torch.full((_TRACK_CHUNK, DIM), 0.5, device=device), generator=g)
        _TORCH_TRACKS[key] = t
    return t


def make_image_torch(image_id: int, n: int, device, *, seed: int = 1234, shared_frac: float = 0.4,
                     track_step: int = 256, noise: float = 0.12):
    """uint8 [n, 128] CUDA tensor: synthetic image ``image_id`` (see make_image for the model)."""
    import torch
    n = int(n)
    if n == 0:
        return torch.empty((0, DIM), dtype=torch.uint8, device=device)
    g = torch.Generator(device=device)
    g.manual_seed((seed * 1000003 + int(image_id)) * 2)
    n_sh = int(n * shared_frac)
    parts = []
    if n_sh:
        first = int(image_id) * track_step
        c0, c1 = first // _TRACK_CHUNK, (first + n_sh - 1) // _TRACK_CHUNK
        base = torch.cat([_track_chunk_torch(c, seed, device) for c in range(c0, c1 + 1)])
        base = base[first - c0 * _TRACK_CHUNK: first - c0 * _TRACK_CHUNK + n_sh]
        parts.append(base + noise * torch.randn(base.shape, device=device, generator=g))
    if n - n_sh:
        parts.append(torch._standard_gamma(torch.full((n - n_sh, DIM), 0.5, device=device), generator=g))
    v = torch.cat(parts).clamp_(min=0.0)
    nrm = torch.linalg.vector_norm(v, dim=1, keepdim=True)
    nrm[nrm == 0] = 1.0
    q = torch.round(v * (512.0 / nrm)).clamp_(0, 255).to(torch.uint8)
    return q[torch.randperm(n, device=device, generator=g)].contiguous()
