"""CPU oracle for the embedding-set distance path — TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference's algorithm (SonyCSLParis/audio-metrics
v1.0.4, ``src/audio_metrics/{data.py,metrics/*.py}``); every function cites the
reference lines it follows.  It exists to check the CUDA path and to provide the
CPU baseline timing; nothing in ``audio_metrics_b200`` imports it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.

Pinning: the reference's own tests hold no golden values for FAD / KD / PRDC /
APA (SURVEY.md §8c — only streaming == single-shot statistics is asserted), so
the oracle is pinned against outputs of the reference itself: the unmodified
reference modules were imported in the build container (``/root/reference``,
with stub modules for five absent audio-only dependencies) and run on seeded
inputs by ``tests/golden/make_golden.py``; the results are committed under
``tests/golden/`` and ``tests/test_oracle.py`` checks every oracle function
against them.  The arithmetic itself lives in third-party libraries (torch
2.11.0 CPU, numpy 2.3.5 in this image — the reference pins neither): the
restatements below follow the published algorithms of ``torch.cov``,
``torch.cdist`` (matmul mode), ``torch.kthvalue`` and ``torch.linalg.eigvals``.
"""
from .stats import batch_stats, chan_merge, StreamingStats  # noqa: F401
from .fad import frechet_from_stats, frechet_sqrtm  # noqa: F401
from .kd import kernel_distance, draw_subset_indices, mmd2, mmd2_unbiased, polynomial_kernel, kd_subset_size  # noqa: F401
from .prdc import cdist_mm, nearest_neighbour_distances, prdc, prdc_counts, prdc_counts_chunked, prdc_bracket  # noqa: F401
from .apa import apa, apa_from_fads  # noqa: F401
