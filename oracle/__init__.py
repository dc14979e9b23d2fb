"""TEST INFRASTRUCTURE ONLY -- float64 NumPy/SciPy restatement of the GP hot path.

Nothing in ``pybo_b200`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs use it, and there only as the checker / the timed CPU arm.

PARITY UNPINNED: the reference (mwhoffman/pybo) keeps its GP arithmetic in the
un-vendored, un-pinned dependency ``reggie`` (``requirements.txt:8``) and ships
no tests, fixtures or golden vectors.  The oracle therefore
follows SURVEY.md section 8a-math (textbook GP regression) and the reference's
own call sites, and is pinned instead against independent implementations
(scikit-learn GPs, scipy.stats, mpmath, torch autograd) by
``tests/test_oracle.py`` and by the committed vectors in ``tests/golden/``.
"""

from .gp_oracle import (  # noqa: F401
    KERNELS,
    GPOracle,
    MixtureOracle,
    FourierSampleOracle,
    thompson_batch_oracle,
    predict_fast,
    kernel_matrix,
    kernel_gradx,
    ucb_beta,
    ucb_index,
    ei_from_moments,
    pi_from_moments,
)
