"""The sobolld sampler's table: data/sobolld/initIrreducibleGF3.dat of the reference (include/wt/sampler/sobolld/irreducible_gf3.hpp:124-158).

The reference ships the table through Git-LFS (a 131-byte pointer stub in /root/reference: SURVEY.md 8c); its origin is the
Quad-Optimized-LDS project (optimised initial direction numbers).  `load_table()` parses the real file when one is available;
`default_table()` builds a structurally identical stand-in -- monic irreducible polynomials over GF(3) in increasing degree and
deterministic initial direction numbers with a non-zero units digit, so every generator matrix is upper triangular with an invertible
diagonal and each dimension is a (0,11,1)-net in base 3 -- clearly NOT the optimised numbers.  The device and oracle arithmetic is
table-agnostic and bit-exact for either.
"""
import ctypes as C
import itertools
import os

from . import _abi as A

ENTRIES, DIMS, DIGITS = 48, 47, 11


def _poly_mod(a, b):
    """remainder of a by monic b over GF(3); coefficient lists, little endian."""
    a = list(a)
    while len(a) >= len(b):
        c = a[-1]
        if c:
            for i in range(len(b)):
                a[len(a) - len(b) + i] = (a[len(a) - len(b) + i] - c * b[i]) % 3
        a.pop()
    return a


def irreducible_polynomials(count):
    """The first `count` monic irreducible polynomials over GF(3), by degree then by value; x itself is skipped."""
    found, deg = [], 1
    while len(found) < count:
        for low in itertools.product(range(3), repeat=deg):
            p = list(reversed(low)) + [1]          # little endian, monic; iterate in increasing numeric value
            if deg == 1 and p[0] == 0:
                continue
            ok = True
            for dd in range(1, deg // 2 + 1):
                for ql in itertools.product(range(3), repeat=dd):
                    q = list(ql) + [1]
                    if not any(_poly_mod(p, q)):
                        ok = False; break
                if not ok: break
            if ok:
                found.append(p)
                if len(found) == count: break
        deg += 1
    return found


def default_table():
    """48 entries (d, sj, aj, mk[0..sj)); entry 0 is the one the reference skips."""
    polys = irreducible_polynomials(ENTRIES)
    out = []
    state = 0x9E3779B97F4A7C15
    for d, p in enumerate(polys, start=1):
        sj = len(p) - 1
        aj = sum(c * 3 ** i for i, c in enumerate(p))
        mk = []
        for i in range(sj):
            state = (state * 6364136223846793005 + 1442695040888963407) & (2 ** 64 - 1)
            m = 1 + (state >> 33) % (3 ** (i + 1) - 1)          # in [1, 3^(i+1))
            if m % 3 == 0: m -= 1                                # units digit non-zero: invertible diagonal
            mk.append(m)
        out.append((d, sj, aj, mk))
    return out


def load_table(path):
    """Parses initIrreducibleGF3.dat as irreducible_gf3_t::load_mk does (lines starting with 'd' skipped, first 48 kept)."""
    entries = []
    with open(path) as f:
        for line in f:
            if line.startswith("d") or not line.strip():
                continue
            v = [int(x) for x in line.split()]
            entries.append((v[0], v[1], v[2], v[3:3 + v[1]]))
            if len(entries) == ENTRIES: break
    if len(entries) < ENTRIES:
        raise RuntimeError(f'Underflow in "{path}"')
    return entries


def is_lfs_stub(path):
    try:
        with open(path, "rb") as f:
            return f.read(40).startswith(b"version https://git-lfs")
    except OSError:
        return True


def table_for(path=None):
    """The real table when `path` holds one, else the stand-in."""
    if path and os.path.exists(path) and not is_lfs_stub(path):
        return load_table(path)
    return default_table()


def to_abi(entries):
    arr = (A.SobolEntry * ENTRIES)()
    for i, (d, sj, aj, mk) in enumerate(entries):
        arr[i].d, arr[i].sj, arr[i].aj = d, sj, aj
        for k, m in enumerate(mk): arr[i].mk[k] = m
    return arr


def host_matrices(entries):
    """Row masks the product library derives from the table (wthost_sobol_tables): (ones, twos) as [47][11] lists."""
    arr = to_abi(entries)
    ones = (C.c_uint16 * (DIMS * DIGITS))(); twos = (C.c_uint16 * (DIMS * DIGITS))()
    A.check_host(A.host_lib().wthost_sobol_tables(arr, ones, twos), "wthost_sobol_tables")
    return [list(ones[d * DIGITS:(d + 1) * DIGITS]) for d in range(DIMS)], [list(twos[d * DIGITS:(d + 1) * DIGITS]) for d in range(DIMS)]
