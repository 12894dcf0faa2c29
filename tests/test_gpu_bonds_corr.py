"""GPU: -J random bonds and -c two-point correlation against the oracle and the reference golden."""
import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("X,Y,prob,xsl,ysl", [(2048, 32, 0.25, 0, 0), (4096, 64, 0.5, 0, 0), (4096, 64, 0.3, 2048, 32),
                                              (2048, 16, 0.0, 0, 0), (2048, 16, 1.0, 0, 0)])
def test_bonds_match_oracle(ig, oracle, X, Y, prob, xsl, ysl):
    T = 1.5
    with ig.IsingEngine(X, Y, temp=T, j_prob=prob, xsl=xsl, ysl=ysl) as eng:
        ora = oracle.OracleLattice(X, Y, xsl=xsl, ysl=ysl)
        eng.init_lattice(); ora.init(); ora.init_bonds(prob)
        rnd, mir = eng.export_bonds_ref_layout()

        def nib(words):     # uint64 4-bit words -> one nibble per site
            out = np.zeros((Y, X // 2), dtype=np.uint8)
            for k in range(16):
                out[:, k::16] = ((words >> np.uint64(4 * k)) & np.uint64(0xF)).astype(np.uint8)
            return out
        assert np.array_equal(nib(rnd), ora.bonds(0)), "random (hamiltInitB_k) buffer"
        assert np.array_equal(nib(mir), ora.bonds(1)), "mirror (hamiltInitW_k) buffer"
        for it in range(1, 4):
            for color in (0, 1):
                eng.update_color(color, it); eng.sync()
                ora.update_color(color, it, T)
                gb, gw = eng.export_bits()
                assert np.array_equal(gb, ora.export_bits(0)) and np.array_equal(gw, ora.export_bits(1))


@pytest.mark.parametrize("X,Y,xsl,ysl", [(4096, 256, 0, 0), (2048, 128, 0, 0), (6144, 160, 0, 0), (4096, 256, 2048, 64)])
def test_correlation_matches_oracle(ig, oracle, X, Y, xsl, ysl):
    T = 2.269185
    with ig.IsingEngine(X, Y, temp=T, xsl=xsl, ysl=ysl) as eng:
        ora = oracle.OracleLattice(X, Y, xsl=xsl, ysl=ysl)
        eng.init_lattice(); ora.init()
        eng.sweeps(6); ora.sweeps(6, T)
        assert np.array_equal(eng.correlation(128), ora.correlation(128))     # exact: integer sums / 2N
        # column 1 is the nearest-neighbour correlation: energy per site = -2*c1
        assert abs(eng.energy_per_site() + 2.0 * eng.correlation(1)[0]) < 1e-12


def test_correlation_multi_device(ig, oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    X, Y, T = 4096, 64, 2.0
    with ig.IsingEngine(X, Y, ndev=2, temp=T) as eng:
        ora = oracle.OracleLattice(X, 2 * Y)
        eng.init_lattice(); ora.init()
        eng.sweeps(3); ora.sweeps(3, T)
        assert np.array_equal(eng.correlation(128), ora.correlation(128))
