"""CPU checks of phylo2vec_b200/csrc/p2v_float_text.cuh (the float64 <-> text routines the Newick-with-lengths
and CSV kernels use), compiled as plain C++ (tests/float_text_host.cpp) so that millions of values are checked
without a GPU:

  * f64_to_shortest against Python's repr -- the same shortest round-trip digits -- laid out by the Rust ryu
    crate's rules (oracle.ryu_format), the format the reference writes (phylo2vec/src/matrix/convert.rs:37-57);
  * parse_f64 against Python's float() -- correctly rounded like Rust's str::parse::<f64>
    (phylo2vec/src/newick/mod.rs:205, 233)."""

import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import p2v_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "float_text_host.cpp")
HDR = os.path.join(ROOT, "phylo2vec_b200", "csrc", "p2v_float_text.cuh")
SO = os.path.join(ROOT, "tests", "_build", "libfloat_text_host.so")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC], check=True)
    return ctypes.CDLL(SO)


def fmt_many(lib, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    buf = np.zeros((x.size, 32), dtype=np.uint8)
    lens = np.zeros(x.size, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)       # noqa: E731
    lib.ft_format_many(p(x), ctypes.c_long(x.size), p(buf), p(lens))
    raw = buf.tobytes()
    return [raw[32 * i:32 * i + lens[i]].decode() for i in range(x.size)]


def parse_many(lib, strings):
    stride = max(len(s) for s in strings) + 1
    buf = np.zeros((len(strings), stride), dtype=np.uint8)
    lens = np.array([len(s) for s in strings], dtype=np.int32)
    for i, s in enumerate(strings):
        buf[i, :len(s)] = np.frombuffer(s.encode(), dtype=np.uint8)
    out = np.zeros(len(strings))
    rc = np.zeros(len(strings), dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)       # noqa: E731
    lib.ft_parse_many(p(buf), p(lens), ctypes.c_long(len(strings)), stride, p(out), p(rc))
    return out, rc


def interesting_doubles(rng, n):
    bits = rng.integers(0, 1 << 64, size=n, dtype=np.uint64)             # every exponent, every mantissa
    x = bits.view(np.float64)
    x = x[np.isfinite(x)]
    short = np.round(rng.random(n // 4) * 10.0 ** rng.integers(-8, 9, n // 4), 6)        # "human" values
    unit = rng.random(n // 4)                                             # what sample_matrix draws
    edges = []
    for e in range(-1074, 1024, 7):
        base = 2.0 ** e
        edges += [base, np.nextafter(base, 0.0), np.nextafter(base, np.inf)]
    for e in range(-323, 309):
        t = float(f"1e{e}")
        edges += [t, np.nextafter(t, 0.0), np.nextafter(t, np.inf), 9.5 * t, 5.0 * t]
    edges += [5e-324, 2.2250738585072014e-308, 2.225073858507201e-308, 1.7976931348623157e308, 9007199254740993.0,
              0.3, 0.1 + 0.2, 1 / 3, 123456.789e3, 1e21, 1e22, 1e23, 4.35, 8.41e21, 2.0 ** 63, 1.0, 0.5]
    return np.concatenate([x, short, unit, np.array(edges), -np.array(edges[:50])])


def test_shortest_formatting_matches_repr_digits_and_ryu_layout(lib):
    rng = np.random.Generator(np.random.Philox(2024))
    x = interesting_doubles(rng, 400_000)
    got = fmt_many(lib, x)
    for xi, g in zip(x.tolist(), got):
        assert g == O.ryu_format(xi), (xi, g)
        assert float(g) == xi                         # and it reads back to the same double
    assert fmt_many(lib, [0.0, -0.0, np.nan, np.inf, -np.inf]) == ["0.0", "-0.0", "NaN", "inf", "-inf"]
    # the reference's own strings (phylo2vec/src/matrix/convert.rs:131-205)
    ref = {0.30118185: "0.30118185", 0.059750594: "0.059750594", 0.0017238181: "0.0017238181", 3.12: "3.12",
           1.0: "1.0", 3.0: "3.0", 0.2021168: "0.2021168"}
    assert fmt_many(lib, list(ref)) == list(ref.values())


def test_parse_is_correctly_rounded(lib):
    rng = np.random.Generator(np.random.Philox(7))
    x = interesting_doubles(rng, 200_000)
    strings = [repr(v) for v in x.tolist()]                                        # shortest strings
    strings += [O.ryu_format(v) for v in x[:50_000].tolist()]                      # the layout the reference writes
    strings += [f"{v:.18e}" for v in x[:50_000].tolist()]                          # numpy.savetxt's layout (io/writer.py:45)
    # random literals of up to 19 digits with exponents across (and beyond) the double range: halfway cases and all
    for _ in range(150_000):
        nd = int(rng.integers(1, 20))
        digs = "".join(map(str, rng.integers(0, 10, nd)))
        e = int(rng.integers(-345, 311))
        form = int(rng.integers(0, 4))
        if form == 0:
            strings.append(f"{digs}e{e}")
        elif form == 1:
            cut = int(rng.integers(0, nd + 1))
            strings.append(f"{digs[:cut]}.{digs[cut:]}E{e:+d}" if nd > 0 and (cut or nd - cut) else f"{digs}e{e}")
        elif form == 2:
            strings.append(("-" if rng.random() < 0.5 else "+") + digs[:1] + "." + digs[1:])
        else:
            strings.append("0." + "0" * int(rng.integers(0, 30)) + digs)
    # exact halfway points between neighbouring doubles (ties-to-even), written out in full when short enough
    for m in rng.integers(1 << 52, 1 << 53, 3000).tolist():
        strings.append(str(2 * m + 1) + "e0" if False else str(2 * m + 1))
        strings.append(str((2 * m + 1) * 5) + "e-1")
    strings = [s for s in strings if s not in (".", "")]
    out, rc = parse_many(lib, strings)
    for s, v, r in zip(strings, out.tolist(), rc.tolist()):
        want = float(s)
        assert r == 0, s
        assert v == want or (v != v and want != want), (s, v, want)
        assert np.signbit(v) == np.signbit(want), s


def test_parse_grammar_and_long_mantissas(lib):
    good = {"1": 1.0, "1.": 1.0, ".5": 0.5, "+.5e1": 5.0, "1e0": 1.0, "1E+2": 100.0, "inf": np.inf, "-INF": -np.inf,
            "Infinity": np.inf, "0e9999": 0.0, "1e-9999": 0.0, "1e9999": np.inf, "00012.500": 12.5, "-0": -0.0}
    out, rc = parse_many(lib, list(good))
    assert not rc.any() and out.tolist() == list(good.values())
    out, rc = parse_many(lib, ["nan", "NaN"])
    assert not rc.any() and np.isnan(out).all()
    bad = ["", ".", "e5", "1e", "1e+", "--1", "1..2", " 1", "1 ", "1f", "0x10", "infin", "1,5", "+", "1e5.0"]
    _, rc = parse_many(lib, [b if b else " " for b in bad])
    assert (rc == 1).all()
    # more than 19 significant digits: decided whenever the 19-digit cut and the cut + 1 agree, and then exact
    rng = np.random.Generator(np.random.Philox(99))
    strings = []
    for _ in range(50_000):
        nd = int(rng.integers(20, 40))
        strings.append("".join(map(str, rng.integers(0, 10, nd))) + f"e{int(rng.integers(-330, 280))}")
    strings += [f"{v:.25e}" for v in rng.random(20_000).tolist()]
    out, rc = parse_many(lib, strings)
    undecided = 0
    for s, v, r in zip(strings, out.tolist(), rc.tolist()):
        if r == 2:
            undecided += 1
        else:
            assert r == 0 and v == float(s), s
    assert undecided <= len(strings) // 400           # about 2^-10 of such literals go on to parse_f64_exact (below)
    # the classic hard case sits exactly on a boundary: reported, never mis-rounded
    _, rc = parse_many(lib, ["9007199254740993.0000000000000000000000001"])
    assert rc[0] in (0, 2)


def parse_full_many(lib, strings, exact_only=False):
    """parse_f64, then parse_f64_exact for what it left undecided -- what the kernels do."""
    stride = max(len(s) for s in strings) + 1
    buf = np.zeros((len(strings), stride), dtype=np.uint8)
    lens = np.array([len(s) for s in strings], dtype=np.int32)
    for i, s in enumerate(strings):
        buf[i, :len(s)] = np.frombuffer(s.encode(), dtype=np.uint8)
    out = np.zeros(len(strings))
    rc = np.zeros(len(strings), dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)       # noqa: E731
    und = ctypes.c_long(0)
    if exact_only:
        lib.ft_parse_exact_many(p(buf), p(lens), ctypes.c_long(len(strings)), stride, p(out), p(rc))
    else:
        lib.ft_parse_full_many(p(buf), p(lens), ctypes.c_long(len(strings)), stride, p(out), p(rc), ctypes.byref(und))
    return out, rc, und.value


def hard_long_literals(rng, n_doubles, n_random):
    """Literals of more than 19 digits: the exact midpoints of neighbouring doubles written out in full (ties go to
    the even mantissa), the same one unit above and below in a far digit, and random long digit strings."""
    from decimal import Decimal, getcontext
    getcontext().prec = 1200
    out = []
    xs = rng.integers(0, 1 << 63, n_doubles, dtype=np.uint64).view(np.float64)
    xs = np.concatenate([xs[np.isfinite(xs) & (xs != 0)], [5e-324, 2.2250738585072014e-308, 2.225073858507201e-308,
                                                          1.0, 0.1, 9007199254740992.0, 1e23, 8.98846567431158e307]])
    for d in xs.tolist():
        d2 = float(np.nextafter(d, np.inf))
        if not np.isfinite(d2):
            continue
        m = (Decimal(d) + Decimal(d2)) / 2
        s = format(m, "f") if 1e-5 < abs(m) < 1e20 else format(m, "e").replace("e+", "e")
        mant, _, ex = s.partition("e")
        tail = "e" + ex if ex else ""
        out.append(s)                                         # the tie itself
        out.append(mant + "1" + tail)                         # just above
        md = list(mant)
        j = len(md) - 1
        while j >= 0 and md[j] in ".0":
            j -= 1
        if j >= 0 and md[j].isdigit():
            md[j] = str(int(md[j]) - 1)
            out.append("".join(md) + "99999" + tail)          # just below
    for _ in range(n_random):
        nd = int(rng.integers(20, 60))
        digs = ("".join(map(str, rng.integers(0, 10, nd)))).lstrip("0") or "1"
        cut = int(rng.integers(0, len(digs) + 1))
        e = int(rng.integers(-340, 311))
        out.append(digs[:cut] + "." + digs[cut:] + (f"e{e}" if rng.random() < 0.7 else ""))
    return out


def test_long_literals_are_decided_exactly(lib):
    rng = np.random.Generator(np.random.Philox(5))
    strings = hard_long_literals(rng, 3000, 40_000)
    want = np.array([float(s) for s in strings])
    out, rc, undecided = parse_full_many(lib, strings)
    assert undecided > 2000                               # the ties and their neighbours do take the slow path
    assert not rc.any()
    assert np.array_equal(out, want) and np.array_equal(np.signbit(out), np.signbit(want))
    # the exact routine alone agrees with the fast one wherever that decides
    out2, rc2, _ = parse_full_many(lib, strings, exact_only=True)
    assert not rc2.any() and np.array_equal(out2, want)
    # 800 significant digits are kept; anything beyond only matters as "not zero"
    one = "1." + "0" * 900 + "1"
    tie = format(__import__("decimal").Decimal(1.0) + __import__("decimal").Decimal(2.0) ** -53, "f")
    out3, rc3, _ = parse_full_many(lib, [one, tie, tie + "0" * 900 + "1", tie[:-1] + "4" + "9" * 900])
    assert not rc3.any() and out3.tolist() == [1.0, 1.0, float(np.nextafter(1.0, 2.0)), 1.0]
