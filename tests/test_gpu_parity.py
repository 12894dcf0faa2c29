"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit for bit.

Mirrors how one would test the reference: run `cuIsing -x X -y Y -n N -t T`, compare the
dumped lattice / spin counts (optimized/main.cu:1140-1209, :1746-1749, :1873-1876).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

T15 = 1.5
TC = 2.269185


def _compare(eng, ora, what):
    gb, gw = eng.export_bits()
    ob, ow = ora.export_bits(0), ora.export_bits(1)
    assert np.array_equal(gb, ob), "%s: black plane differs in %d words" % (what, int((gb != ob).sum()))
    assert np.array_equal(gw, ow), "%s: white plane differs in %d words" % (what, int((gw != ow).sum()))
    assert eng.count_spins() == ora.count()


@pytest.mark.parametrize("X,Y", [(2048, 16), (2048, 32), (4096, 48), (6144, 16)])
@pytest.mark.parametrize("temp", [T15, TC])
def test_small_lattice_every_half_sweep(ig, oracle, X, Y, temp):
    """Minimum legal lattice and a few ragged shapes: identical after init and after each colour."""
    with ig.IsingEngine(X, Y, temp=temp) as eng:
        ora = oracle.OracleLattice(X, Y)
        eng.init_lattice(); ora.init()
        _compare(eng, ora, "init")
        for it in range(1, 4):
            for color in (0, 1):
                eng.update_color(color, it); eng.sync()
                ora.update_color(color, it, temp)
                _compare(eng, ora, "it %d colour %d" % (it, color))


def test_known_small_counts(ig):
    """SURVEY appendix A3 (oracle-derived, cross-checked with the reference binary on B200)."""
    with ig.IsingEngine(2048, 16, temp=T15) as eng:
        eng.init_lattice()
        got = [eng.count_spins()]
        for _ in range(4):
            eng.sweeps(1); got.append(eng.count_spins())
    assert got == [(16475, 16293), (16699, 16069), (16776, 15992), (16947, 15821), (17161, 15607)]


@pytest.mark.parametrize("temp", [T15, TC])
def test_4096_square_16_sweeps(ig, oracle, temp):
    with ig.IsingEngine(4096, 4096, temp=temp) as eng:
        ora = oracle.OracleLattice(4096, 4096)
        eng.init_lattice(); ora.init()
        eng.sweeps(16); ora.sweeps(16, temp)
        _compare(eng, ora, "4096^2 after 16 sweeps")
        e_gpu, e_cpu = eng.energy_per_site(), ora.energy()
        assert abs(e_gpu - e_cpu) < 1e-12


@pytest.mark.parametrize("temp", [0.3, 0.1, 0.0, 50.0, 1e9])
def test_extreme_temperatures_generic_rule(ig, oracle, temp):
    """Thresholds that degenerate to never/always (exp underflow, T<=0 branch main.cu:1689-1693)."""
    with ig.IsingEngine(2048, 32, temp=temp) as eng:
        ora = oracle.OracleLattice(2048, 32)
        eng.init_lattice(); ora.init()
        eng.sweeps(3); ora.sweeps(3, temp)
        _compare(eng, ora, "T=%g" % temp)


def test_seed_and_iteration_numbering(ig, oracle):
    """Another seed with a high word, and sweeps issued in two calls continue the `it` sequence."""
    seed = 0x1234567890ABCDEF
    with ig.IsingEngine(4096, 32, temp=TC, seed=seed) as eng:
        ora = oracle.OracleLattice(4096, 32, seed=seed)
        eng.init_lattice(); ora.init()
        eng.sweeps(2); eng.sweeps(3)
        ora.sweeps(5, TC)
        _compare(eng, ora, "seed/iteration")


def test_sublattices(ig, oracle):
    """--xsl/--ysl: independent periodic sub-lattices inside one buffer (main.cu:1423-1463)."""
    with ig.IsingEngine(4096, 64, temp=T15, xsl=2048, ysl=16) as eng:
        ora = oracle.OracleLattice(4096, 64, xsl=2048, ysl=16)
        eng.init_lattice(); ora.init()
        eng.sweeps(4); ora.sweeps(4, T15)
        _compare(eng, ora, "sub-lattices")


def test_ref_layout_roundtrip_and_dump(ig, oracle, tmp_path):
    with ig.IsingEngine(2048, 32, temp=T15) as eng:
        ora = oracle.OracleLattice(2048, 32)
        eng.init_lattice(); ora.init()
        eng.sweeps(2); ora.sweeps(2, T15)
        rb, rw = eng.export_ref_layout()
        assert np.array_equal(rb, ora.export_ref(0)) and np.array_equal(rw, ora.export_ref(1))
        eng.dump_lattice(str(tmp_path / "g_")); ora.dump(str(tmp_path / "o.txt"))
        assert (tmp_path / "g_0.txt").read_bytes() == (tmp_path / "o.txt").read_bytes()
        # import the reference layout into a fresh context and continue: same trajectory
        with ig.IsingEngine(2048, 32, temp=T15) as e2:
            e2.import_ref_layout(rb, rw); e2.it = 2
            e2.sweeps(2); eng.sweeps(2)
            b1, w1 = eng.export_bits(); b2, w2 = e2.export_bits()
            assert np.array_equal(b1, b2) and np.array_equal(w1, w2)


def test_readme_kat1_full_size(ig):
    """optimized/README.md:128-130: 65536x65536, T=1.5, default seed, counts at it 0, 16, 32."""
    with ig.IsingEngine(65536, 65536, temp=T15) as eng:
        eng.init_lattice()
        assert eng.count_spins() == (2147484090, 2147483206)
        eng.sweeps(16)
        assert eng.count_spins() == (2147575418, 2147391878)
        eng.sweeps(16)
        assert eng.count_spins() == (2147641872, 2147325424)


def test_run_from_host_matches_resident_path(ig):
    with ig.IsingEngine(4096, 256, temp=TC) as eng:
        eng.init_lattice()
        b0, w0 = eng.export_bits()
        eng.sweeps(4)
        b1, w1 = eng.export_bits()
        c1 = eng.count_spins()
    with ig.IsingEngine(4096, 256, temp=TC) as e2:
        bo, wo = np.empty_like(b0), np.empty_like(w0)
        c2 = e2.run_from_host(b0, w0, 4, bo, wo)
        assert np.array_equal(bo, b1) and np.array_equal(wo, w1) and c1 == c2


def test_replica_batch_from_host_matches_oracle(ig, oracle):
    """ising_run_batch_from_host: 5 different lattices (even / odd jobs on different contexts)
    uploaded from host, 3 sweeps each, downloaded: every replica equals the oracle run from the
    same start; a NULL second context gives the same result."""
    X, Y, n = 4096, 64, 3
    rng = np.random.default_rng(7)
    starts = [(rng.integers(0, 2**32, size=(Y, X // 64), dtype=np.uint32),
               rng.integers(0, 2**32, size=(Y, X // 64), dtype=np.uint32)) for _ in range(5)]
    want = []
    for b, w in starts:
        ora = oracle.OracleLattice(X, Y)
        ora.import_bits(0, b); ora.import_bits(1, w)
        ora.sweeps(n, TC)
        want.append((ora.export_bits(0), ora.export_bits(1), ora.count()))
        ora.close()
    with ig.IsingEngine(X, Y, temp=TC) as ea, ig.IsingEngine(X, Y, temp=TC) as eb, ig.IsingEngine(X, Y, temp=TC) as ec:
        for other in (eb, None, [eb, ec], []):      # 2 contexts, 1, 3 (ising_run_batch_from_host_n), 1 through _n
            ob = [np.zeros((Y, X // 64), np.uint32) for _ in starts]
            ow = [np.zeros((Y, X // 64), np.uint32) for _ in starts]
            counts = ea.run_batch_from_host(other, [s[0] for s in starts], [s[1] for s in starts], n, ob, ow)
            for k, (wb, ww, wc) in enumerate(want):
                assert np.array_equal(ob[k], wb) and np.array_equal(ow[k], ww), "replica %d differs" % k
                assert counts[k] == wc
        # outputs and counts are optional
        assert ea.run_batch_from_host(eb, [starts[0][0]], [starts[0][1]], n) == [want[0][2]]


def test_checkpoint_continues_the_run(ig, oracle, tmp_path):
    """n1 sweeps, save, load into a fresh context, n2 sweeps == n1 + n2 sweeps (same Philox stream);
    shape / seed mismatches and foreign files are refused."""
    X, Y, n1, n2 = 4096, 48, 3, 4
    ck = str(tmp_path / "run.ckpt")
    with ig.IsingEngine(X, Y, temp=TC) as eng:
        eng.init_lattice(); eng.sweeps(n1)
        eng.save_checkpoint(ck)
    ora = oracle.OracleLattice(X, Y)
    ora.init(); ora.sweeps(n1 + n2, TC)
    with ig.IsingEngine(X, Y, temp=TC) as eng:
        assert eng.load_checkpoint(ck) == n1
        eng.sweeps(n2)
        _compare(eng, ora, "continued run")
    with ig.IsingEngine(X, Y + 16, temp=TC) as eng, pytest.raises(ig.IsingError, match="does not match"):
        eng.load_checkpoint(ck)
    with ig.IsingEngine(X, Y, temp=TC, seed=7) as eng, pytest.raises(ig.IsingError, match="seed"):
        eng.load_checkpoint(ck)
    bad = tmp_path / "bad.ckpt"
    bad.write_bytes(b"not a checkpoint" * 8)
    with ig.IsingEngine(X, Y, temp=TC) as eng, pytest.raises(ig.IsingError, match="not an ising_b200 checkpoint"):
        eng.load_checkpoint(str(bad))
