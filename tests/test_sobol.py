"""sobolld sampler: the oracle's literal restatement of the reference's generator, the product's host-side
generator matrices, and -- on the GPU -- bit-exact numerators from the in-register device generator.

The reference holds no test or golden vector for this sampler and its table is a Git-LFS stub, so the pins are the
properties the construction guarantees for ANY valid table: each dimension of a full scrambled batch is a permutation of the 3^11
numerators (a (0,11,1)-net), consecutive blocks of 3^m points stratify 3^m intervals, and the scrambling is a pure function of
(seed, dimension, batch)."""
import ctypes as C
import os
import numpy as np
import pytest

from wave_tracer_b200 import _abi as A, sobol
import _oracle

NPTS, D, M = 3 ** 11, 47, 11


def oracle_batch(table, seed, batch, n):
    arr = sobol.to_abi(table)
    num = np.zeros(n * D, np.uint32); val = np.zeros(n * D, np.float32)
    got = _oracle.lib().oracle_sobol_batch(arr, seed, batch, n, num.ctypes.data_as(C.POINTER(C.c_uint32)), val.ctypes.data_as(C.POINTER(C.c_float)))
    assert got == n
    return num.reshape(n, D), val.reshape(n, D)


def test_default_table_is_wellformed():
    t = sobol.default_table()
    assert len(t) == 48 and len({aj for _, _, aj, _ in t}) == 48
    for d, sj, aj, mk in t:
        assert 1 <= sj <= 10 and len(mk) == sj and 3 ** sj <= aj < 3 ** (sj + 1) * 1      # monic of degree sj... leading digit 1
        assert aj // 3 ** sj == 1
        for i, m in enumerate(mk): assert 0 < m < 3 ** (i + 1) and m % 3 != 0
    # no linear factor: a polynomial of degree >= 2 without a root in GF(3) (full irreducibility is what sobol._poly_mod trial division checks)
    for d, sj, aj, mk in t:
        if sj < 2: continue
        c = [(aj // 3 ** i) % 3 for i in range(sj + 1)]
        assert all(sum(ci * x ** i for i, ci in enumerate(c)) % 3 != 0 for x in range(3)), (d, c)


def test_dat_parser_roundtrip(tmp_path):
    t = sobol.default_table()
    p = tmp_path / "initIrreducibleGF3.dat"
    p.write_text("d sj aj mk\n" + "\n".join(f"{d} {sj} {aj} " + " ".join(map(str, mk)) for d, sj, aj, mk in t) + "\n")
    assert sobol.load_table(str(p)) == t
    assert sobol.table_for(str(p)) == t
    stub = tmp_path / "stub.dat"; stub.write_text("version https://git-lfs.github.com/spec/v1\noid sha256:00\nsize 922\n")
    assert sobol.table_for(str(stub)) == t         # pointer stub -> stand-in table


def test_host_matrices_match_oracle_gen_mat():
    """wthost_sobol_tables (product, digit-vector recurrence) == generate_mkgf3 + gen_mat (oracle, integer encode/decode as the reference)."""
    t = sobol.default_table()
    ones, twos = sobol.host_matrices(t)
    mat = np.zeros(D * M * M, np.int32)
    assert _oracle.lib().oracle_sobol_matrices(sobol.to_abi(t), mat.ctypes.data_as(C.POINTER(C.c_int32))) == 0
    mat = mat.reshape(D, M, M)
    for d in range(D):
        assert np.all(np.tril(mat[d], -1) == 0) and np.all(np.diag(mat[d]) != 0)          # upper triangular, invertible
        for j in range(M):
            row = mat[d, M - 1 - j]
            assert ones[d][j] == sum(1 << c for c in range(M) if row[c] == 1)
            assert twos[d][j] == sum(1 << c for c in range(M) if row[c] == 2)


def test_oracle_batch_is_a_scrambled_net():
    num, val = oracle_batch(sobol.default_table(), 0x5EED, 0, NPTS)
    assert val.min() >= 0 and val.max() < 1
    assert np.array_equal(val, (num.astype(np.float32) / np.float32(NPTS)))
    for d in (0, 1, 7, 23, 46):
        assert np.array_equal(np.sort(num[:, d]), np.arange(NPTS, dtype=np.uint32))        # permutation of all numerators
        for m in (1, 3, 6):                                                                # every aligned block of 3^m points stratifies 3^m intervals
            blk = num[:3 ** m * 5, d].reshape(5, 3 ** m) // 3 ** (M - m)
            assert all(len(set(b)) == 3 ** m for b in blk)
    # different batches / seeds scramble differently, same inputs reproduce
    n2, _ = oracle_batch(sobol.default_table(), 0x5EED, 1, 64); n3, _ = oracle_batch(sobol.default_table(), 0x5EEE, 0, 64); n4, _ = oracle_batch(sobol.default_table(), 0x5EED, 0, 64)
    assert not np.array_equal(n2, num[:64]) and not np.array_equal(n3, num[:64]) and np.array_equal(n4, num[:64])


def test_oracle_render_with_sobolld_scene_sampler():
    """The sampler only changes which emitter / wavenumber / sensor-position draws a sample gets: images agree statistically with the uniform sampler."""
    from wave_tracer_b200 import scenes
    from wave_tracer_b200.scene import Sobolld
    sc = scenes.cornell_like(res=24, spp=8, max_depth=4, n_sphere=6); b0 = sc.build()
    sc.sampler = Sobolld(); b1 = sc.build()
    assert b1.sampler == A.SAMPLER_SOBOLLD and bool(b1.desc.sobol_table)
    blk0, _, st0 = _oracle.render(b0, spp=8); blk1, _, st1 = _oracle.render(b1, spp=8)
    assert st0["samples"] == st1["samples"]
    m0 = blk0[..., 0].sum() / blk0[..., 1].sum(); m1 = blk1[..., 0].sum() / blk1[..., 1].sum()
    assert m0 > 0 and abs(m0 - m1) / m0 < 0.15
    assert not np.array_equal(blk0, blk1)
    # stratification of the sensor-position draws: per-pixel weight sums (Gaussian filter of the in-pixel offsets) vary less than with uniform draws
    blk1b, _, _ = _oracle.render(b1, spp=8)
    assert np.array_equal(blk1, blk1b)          # deterministic


@pytest.mark.gpu
def test_device_sobol_bit_exact_vs_oracle():
    from wave_tracer_b200 import scenes, GpuScene
    from wave_tracer_b200.scene import Sobolld
    sc = scenes.cornell_like(res=16, spp=4, n_sphere=4); sc.sampler = Sobolld()
    built = sc.build(); gs = GpuScene(built, 0)
    t = sc.sampler.table
    for seed, batch, first, n in ((0x5EED, 0, 0, 4096), (0x5EED, 0, NPTS - 300, 300), (0x1234567890ABCDEF, 3, 1000, 2000), (7, 2 ** 33 + 5, 50000, 512)):
        num = np.zeros(n * D, np.uint32); val = np.zeros(n * D, np.float32)
        A.check(A.lib().wtgpu_debug_sobol(gs.handle, seed, batch * NPTS + first, n, num.ctypes.data_as(C.POINTER(C.c_uint32)), val.ctypes.data_as(C.POINTER(C.c_float))), "wtgpu_debug_sobol")
        onum, oval = oracle_batch(t, seed, batch, first + n)
        assert np.array_equal(num.reshape(n, D), onum[first:]), "sobol numerators differ from the reference arithmetic"
        assert np.array_equal(val.reshape(n, D).view(np.uint32), oval[first:].view(np.uint32))
    # a range crossing a batch boundary: the second part is the start of the next batch
    n = 64; num = np.zeros(n * D, np.uint32); val = np.zeros(n * D, np.float32)
    A.check(A.lib().wtgpu_debug_sobol(gs.handle, 0x5EED, NPTS - 32, n, num.ctypes.data_as(C.POINTER(C.c_uint32)), val.ctypes.data_as(C.POINTER(C.c_float))), "wtgpu_debug_sobol")
    o0, _ = oracle_batch(t, 0x5EED, 0, NPTS); o1, _ = oracle_batch(t, 0x5EED, 1, 32)
    assert np.array_equal(num.reshape(n, D), np.concatenate([o0[-32:], o1]))
    gs.close()


@pytest.mark.gpu
@pytest.mark.parametrize("integrator", ["plt_path", "plt_bdpt"])
def test_gpu_render_with_sobolld_matches_oracle(integrator):
    from wave_tracer_b200 import scenes, render
    from wave_tracer_b200.scene import Sobolld
    sc = scenes.cornell_like(res=32, spp=8, max_depth=5, n_sphere=6, integrator=integrator); sc.sampler = Sobolld()
    built = sc.build()
    blk, lgt, st = render(built, spp=8, device=0)
    oblk, olgt, ost = _oracle.render(built, spp=8)
    assert st["samples"] == ost["samples"]
    num = np.linalg.norm(blk.astype(np.float64) - oblk); den = np.linalg.norm(oblk)
    assert den > 0 and num / den <= 5e-3, num / den         # rel-L2 tolerance of the uniform-sampler parity tests (f32 path math + f32 film atomics vs f64 film)


# ------------------------------------------------------------------------------------------------ pinned against the reference's own code
REF_SOBOL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_sobol.so")


def _write_dat(path, table):
    """initIrreducibleGF3.dat layout (irreducible_gf3.hpp:124-158): a header line starting with 'd', then one line `d sj aj mk...` per entry."""
    with open(path, "w") as f:
        f.write("d sj aj mk\n")
        for d, sj, aj, mk in table:
            f.write(" ".join(str(v) for v in [d, sj, aj] + list(mk)) + "\n")


@pytest.mark.skipif(not os.path.exists(REF_SOBOL), reason="oracle/_ref/libref_sobol.so is built from /root/reference (this container only)")
def test_oracle_sobol_equals_the_reference_code(tmp_path):
    """The restatement (oracle/ot_sobol.h) against the REFERENCE'S OWN sobolld headers, compiled unmodified into oracle/_ref/libref_sobol.so
    (oracle/ref_sobol.cpp + two shim headers): the table file goes through the reference's parser, generator matrices and 3^8 points x 47
    dimensions are compared BIT FOR BIT, for the seeds of our batch contract and for arbitrary ones.  Together with the golden fixture
    (tests/golden: oracle numerators for seed 0x5EED, batch 0) and the GPU test of the device generator against the oracle, this pins the
    Sobol index / digit / scramble math end to end."""
    R = C.CDLL(REF_SOBOL)
    R.ref_sobol_points.argtypes = [C.c_char_p, C.POINTER(C.c_uint64), C.c_uint32, C.POINTER(C.c_float)]
    R.ref_sobol_matrices.argtypes = [C.c_char_p, C.POINTER(C.c_int32)]
    L = _oracle.lib()
    table = sobol.default_table()
    dat = str(tmp_path / "initIrreducibleGF3.dat"); _write_dat(dat, table)
    assert sobol.load_table(dat) == [(d, sj, aj, list(mk)) for d, sj, aj, mk in table]
    abi = sobol.to_abi(table)
    m_ref = np.zeros(47 * 11 * 11, np.int32); m_or = np.zeros_like(m_ref)
    assert R.ref_sobol_matrices(dat.encode(), m_ref.ctypes.data_as(C.POINTER(C.c_int32))) == 0
    assert L.oracle_sobol_matrices(abi, m_or.ctypes.data_as(C.POINTER(C.c_int32))) == 0
    assert np.array_equal(m_ref, m_or) and m_ref.any()
    n = 3 ** 8
    rng = np.random.default_rng(11)
    seed_sets = []
    s0 = np.zeros(47, np.uint64); L.oracle_sobol_seeds(0x5EED, 0, s0.ctypes.data_as(C.POINTER(C.c_uint64))); seed_sets.append(s0)
    s1 = np.zeros(47, np.uint64); L.oracle_sobol_seeds(0x5EED, 3, s1.ctypes.data_as(C.POINTER(C.c_uint64))); seed_sets.append(s1)
    seed_sets.append(rng.integers(0, 2 ** 32, 47, dtype=np.uint64))          # what the reference feeds it: uniform_int_distribution<unsigned>
    seed_sets.append(rng.integers(0, 2 ** 63, 47, dtype=np.uint64))
    for seeds in seed_sets:
        seeds = np.ascontiguousarray(seeds, np.uint64)
        v_ref = np.zeros(n * 47, np.float32); v_or = np.zeros(n * 47, np.float32); num = np.zeros(n * 47, np.uint32)
        assert R.ref_sobol_points(dat.encode(), seeds.ctypes.data_as(C.POINTER(C.c_uint64)), n, v_ref.ctypes.data_as(C.POINTER(C.c_float))) == n
        assert L.oracle_sobol_points_with_seeds(abi, seeds.ctypes.data_as(C.POINTER(C.c_uint64)), n, num.ctypes.data_as(C.POINTER(C.c_uint32)), v_or.ctypes.data_as(C.POINTER(C.c_float))) == n
        assert np.array_equal(v_ref.view(np.uint32), v_or.view(np.uint32))                  # bit-identical floats
        assert np.array_equal(np.float32(num) / np.float32(3 ** 11), v_ref)                 # value = numerator / 3^M (integer3.hpp:48-50)
        assert len(np.unique(v_ref.reshape(n, 47)[:, 0])) == n                              # a (0,8,1)-net in base 3: all first coordinates distinct
    # the frozen golden numerators (oracle, seed 0x5EED, batch 0) are what the reference code produces for those seeds
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz"))
    g = G["sobol/numerators_seed5EED_batch0_81pts"]
    v_ref = np.zeros(81 * 47, np.float32)
    assert R.ref_sobol_points(dat.encode(), np.ascontiguousarray(seed_sets[0]).ctypes.data_as(C.POINTER(C.c_uint64)), 81, v_ref.ctypes.data_as(C.POINTER(C.c_float))) == 81
    assert np.array_equal(np.float32(g) / np.float32(3 ** 11), v_ref.reshape(81, 47))
