"""Synthetic embedding matrices for the hot path's tests and benchmarks.

The reference feeds ``PerformClusteringWithConstraints`` one fp32 row per image:
a ResNet50 feature concatenated with a label one-hot block
(``/root/reference/internal/embeddings/embeddings.go:166-183``,
``internal/workflow/workflow.go:167-168``).  BASELINE.json measures on synthetic
Gaussian mixtures of that width (D=2048, or 2048+100 for the combined vector).
"""
from __future__ import annotations

import numpy as np

# BASELINE.json configs: name -> (N, D, minSize, maxSize)
CONFIGS = {
    "A": (1_000, 2048, 5, 20),
    "B": (20_000, 2048, 10, 50),
    "C": (100_000, 2048, 20, 200),
    "D": (250_000, 2048, 20, 200),
    "E": (50_000, 2148, 2, 8),
}


def gaussian_mixture(n: int, d: int, min_size: int, max_size: int, seed: int = 20240,
                     sigma: float = 0.3, relu_like: bool = False, out: np.ndarray | None = None) -> np.ndarray:
    """[n x d] fp32, K = n // ((min+max)//2) components, means ~ N(0, I), noise sigma.

    Generated in row blocks so that N=250k x 2048 never needs a float64 temporary
    of the full matrix.  ``relu_like`` maps x -> max(x + 1, 0): non-negative with a
    large common mean, which stresses the Gram identity's cancellation.
    """
    rng = np.random.default_rng(seed)
    k = max(1, n // max(1, (min_size + max_size) // 2))
    means = rng.standard_normal((k, d), dtype=np.float32)
    labels = rng.integers(0, k, size=n)
    x = out if out is not None else np.empty((n, d), dtype=np.float32)
    blk = 8192
    for s in range(0, n, blk):
        e = min(n, s + blk)
        noise = rng.standard_normal((e - s, d), dtype=np.float32)
        noise *= np.float32(sigma)
        noise += means[labels[s:e]]
        if relu_like:
            noise += np.float32(1.0)
            np.maximum(noise, 0, out=noise)
        x[s:e] = noise
    return x


def combined_features(n: int, d_img: int = 2048, n_labels: int = 100, min_size: int = 2, max_size: int = 8,
                      seed: int = 20244, max_ones: int = 10) -> np.ndarray:
    """Config E: image block as ``gaussian_mixture`` + a 0/1 label block with at most
    ``max_ones`` ones per row (Rekognition maxLabels=10, workflow.go:129;
    GenerateLabelVector writes exactly 0 or 1, embeddings.go:166-174)."""
    x = np.zeros((n, d_img + n_labels), dtype=np.float32)
    x[:, :d_img] = gaussian_mixture(n, d_img, min_size, max_size, seed=seed)
    rng = np.random.default_rng(seed + 1)
    counts = rng.integers(0, max_ones + 1, size=n)
    for i in range(n):
        if counts[i]:
            x[i, d_img + rng.choice(n_labels, size=counts[i], replace=False)] = 1.0
    return x


def item_ids(n: int) -> list[str]:
    """ids as the caller builds them, workflow.go:140."""
    return [f"img_{i}" for i in range(n)]
