"""STAGE RULE of the bf16 parity split (tests/parity.py): every kernel of the encoder layer and of the decoder layer, fed the
same-rounding-points oracle's OWN stored operands (float64 arithmetic + the kernels' stores, oracle.kernel_storage_rounding),
reproduces that oracle's stored output within 5e-4 relative L2.  This is the finest granularity at which "the kernels' own
error" is observable: chaining two or more stores lets 1-ulp rounding flips compound (delta -> sqrt(delta x ulp) per store,
saturating at the rounding noise itself, ~1e-3), whatever the kernels do — which is why whole decoder layers sit at
1.3 - 1.9e-3 against the same oracle while each of their kernels is below 2e-4 (scripts/diag_rounding.py prints both)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))

STAGE_TOL = 5e-4


@pytest.mark.parametrize("tile", [128, 256])
def test_every_kernel_matches_the_same_rounding_oracle_on_its_operands(cuda_dev, tile):
    import diag_rounding
    stages = diag_rounding.run(tile)
    assert len(stages) >= 18
    bad = {k: v for k, v in stages.items() if not v < STAGE_TOL}
    assert not bad, bad
