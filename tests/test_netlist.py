"""Netlist front end (bg_netlist_*): hand-derived stamps, the reference's own parser compiled from
/root/reference (oracle/_ref/libref_parser.so) as the bit-exact pin, the reduced formulation against the full MNA
system, error behaviour.  CPU only."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))
import b200gmres as bg  # noqa: E402
from b200gmres import synth  # noqa: E402

bg.build()  # make -C gpu-gmres_b200 when the library is not there yet (host code only is exercised here)

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_parser.so")


def dense(N, which):
    rp, ci, v = getattr(N, which)
    ncol = N.info.n_sources if which == "B" else N.info.size
    return sp.csr_matrix((v, ci, rp), shape=(N.info.size, ncol))


def ref_parse(path):
    """The reference's parser_sub + stampB/stampC/stampG on `path` -> dict (column-compressed matrices as it builds them)."""
    L = C.CDLL(REF_SO)
    saved = os.dup(0)
    null = os.open(os.devnull, os.O_RDONLY)
    os.dup2(null, 0)  # its DC current-source branch polls stdin (src/parser.cpp: "while( !getchar() )")
    try:
        rc = L.refp_parse(str(path).encode())
    finally:
        os.dup2(saved, 0)
        os.close(saved)
        os.close(null)
    assert rc == 0
    counts = (C.c_int * 5)()
    tran = (C.c_double * 2)()
    nnz = (C.c_long * 3)()
    L.refp_info(counts, tran, nnz)
    nn, nl, nv, ni, nports = counts[:]
    size, nsrc = nn + nl + nv, nv + ni
    out = dict(n_nodes=nn, n_l=nl, n_vs=nv, n_is=ni, tstep=tran[0], tstop=tran[1], size=size)
    for k, name in enumerate("GCB"):
        ncol = nsrc if name == "B" else size
        cp, ri, x = np.zeros(ncol + 1, np.int64), np.zeros(nnz[k], np.int64), np.zeros(nnz[k])
        if nnz[k]:
            L.refp_matrix(k, cp.ctypes.data_as(C.c_void_p), ri.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p))
        out[name] = sp.csc_matrix((x, ri, cp), shape=(size, ncol))
    src = []
    for k in range(nsrc):
        m = L.refp_source_len(k)
        t, v = np.zeros(m), np.zeros(m)
        L.refp_source(k, t.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
        src.append((t, v))
    out["sources"] = src
    ports = np.zeros(nports, np.int32)
    L.refp_ports(ports.ctypes.data_as(C.c_void_p))
    out["ports"] = ports
    buf = C.create_string_buffer(256)
    names = []
    for r in range(nn):
        assert L.refp_node_name(r, buf, 256) == 0
        names.append(buf.value.decode())
    out["names"] = names
    npc = L.refp_pulse_current(None, None)
    pt, pv = np.zeros((npc, 5)), np.zeros((npc, 2))
    L.refp_pulse_current(pt.ctypes.data_as(C.c_void_p), pv.ctypes.data_as(C.c_void_p))
    out["pulse_time"], out["pulse_val"] = pt, pv
    return out


def same_matrix(mine, ref):
    """bit-exact: same pattern (explicitly stored entries included), same doubles"""
    a, b = mine.tocoo(), ref.tocsr().tocoo()
    ka, kb = np.lexsort((a.col, a.row)), np.lexsort((b.col, b.row))
    return (a.shape == b.shape and len(a.data) == len(b.data) and np.array_equal(a.row[ka], b.row[kb])
            and np.array_equal(a.col[ka], b.col[kb]) and np.array_equal(a.data[ka], b.data[kb]))


def compare_with_reference(N, R):
    i = N.info
    assert (i.n_nodes, i.n_inductors, i.n_vsources, i.n_isources, i.size) == (R["n_nodes"], R["n_l"], R["n_vs"], R["n_is"], R["size"])
    assert (i.tstep, i.tstop) == (R["tstep"], R["tstop"])
    assert [N.row_name(r) for r in range(i.n_nodes)] == list(R["names"])
    for w in "GCB":
        assert same_matrix(dense(N, w), R[w]), w
        cp, ri, x = N.csc(w)      # the cs_dl arrays themselves, element for element
        ref = R[w].tocsc()
        assert np.array_equal(cp, ref.indptr) and np.array_equal(ri, ref.indices) and np.array_equal(x, ref.data), w
    assert np.array_equal(N.ports, R["ports"])
    for k, ((t, v), (rt, rv)) in enumerate(zip(N.pwl, R["sources"])):
        # the reference trims a PWL table only when the next source of the list is PWL too or at .end
        # (src/parser.cpp:2367-2370, 2858-2867); otherwise it stays padded to a multiple of PWL_SIZE = 800
        assert len(rt) == len(t) or (len(rt) % 800 == 0 and len(rt) > len(t)), k
        assert np.array_equal(t, rt[:len(t)]) and np.array_equal(v, rv[:len(t)]), k


def assert_matches_reference(text, tmp_path, name="case.sp"):
    path = tmp_path / name
    path.write_text(text)
    N = bg.Netlist(path)
    R = ref_parse(path)
    compare_with_reference(N, R)
    return N, R


def load_reference_golden(name):
    """tests/golden/netlist_*.npz (tests/golden/make_frontend_golden.py): what the reference's compiled parser produced"""
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    nn, nl, nv, ni = (int(x) for x in g["counts"])
    R = dict(n_nodes=nn, n_l=nl, n_vs=nv, n_is=ni, size=nn + nl + nv, tstep=float(g["tran"][0]), tstop=float(g["tran"][1]),
             ports=g["ports"], names=[str(x) for x in g["names"]], pulse_time=g["pulse_time"], pulse_val=g["pulse_val"])
    for w in "GCB":
        R[w] = sp.csc_matrix((g[w + "_x"], g[w + "_i"], g[w + "_p"]), shape=tuple(int(x) for x in g[w + "_shape"]))
    ptr = g["src_ptr"]
    R["sources"] = [(g["src_t"][a:b], g["src_v"][a:b]) for a, b in zip(ptr[:-1], ptr[1:])]
    return str(g["text"]), R


@pytest.mark.parametrize("name", ["netlist_grid_pwl", "netlist_grid_pulse", "netlist_mixed"])
def test_equals_committed_reference_output(name):
    """the same comparison against the committed output of the reference's parser (runs without /root/reference)"""
    text, R = load_reference_golden(name)
    N = bg.Netlist(text, text=True)
    compare_with_reference(N, R)
    if len(R["pulse_time"]):      # the reference records the PULSE parameters of current sources only
        cur = N.pulse_src >= N.info.n_vsources
        assert np.array_equal(N.pulse_time[cur], R["pulse_time"]) and np.array_equal(N.pulse_val[cur], R["pulse_val"])


TINY = """* two nodes, an inductor, a pad and a load
R1 a b 2k
R2 b 0 4
R3 a b 1MEG
C1 a 0 3p
C2 a b 1f
L1 b c 5n
V1 c gnd 1.8
I1 a 0 PWL(1n 1m 2n 2m)
.tran 1p 4n
.print tran v(a) v(c) v(0) v(nosuch)
.end
"""


def test_hand_derived_stamps():
    N = bg.Netlist(TINY, text=True)
    i = N.info
    assert (i.n_nodes, i.n_inductors, i.n_vsources, i.n_isources, i.size, i.n_sources, i.dc) == (3, 1, 1, 1, 5, 2, 0)
    assert [N.row_name(r) for r in range(5)] == ["a", "b", "c", "i(L#0)", "i(V#0)"]
    g1, g2, g3 = 1.0 / (2 * 10.0 ** 3), 1.0 / 4, 1.0 / (1 * 10.0 ** 6)
    G = np.zeros((5, 5))
    G[0, 0] = g1 + g3; G[1, 1] = g1 + g2 + g3; G[0, 1] = G[1, 0] = -g1 - g3
    G[3, 1], G[1, 3], G[3, 2], G[2, 3] = -1, 1, 1, -1      # L1 b c: row 3
    G[2, 4], G[4, 2] = 1, -1                                 # V1 c gnd: row 4
    Cm = np.zeros((5, 5))
    Cm[0, 0] = 3 * 10.0 ** -12 + 1 * 10.0 ** -15; Cm[1, 1] = 1 * 10.0 ** -15; Cm[0, 1] = Cm[1, 0] = -1 * 10.0 ** -15
    Cm[3, 3] = 5 * 10.0 ** -9
    B = np.zeros((5, 2))
    B[4, 0] = -1   # V1
    B[0, 1] = -1   # I1 flows out of a
    assert np.array_equal(dense(N, "G").toarray(), G)
    assert np.array_equal(dense(N, "C").toarray(), Cm)
    assert np.array_equal(dense(N, "B").toarray(), B)
    assert (i.tstep, i.tstop) == (1 * 10.0 ** -12, 4 * 10.0 ** -9)
    (tv, vv), (ti, vi) = N.pwl
    assert np.array_equal(tv, [0, i.tstop]) and np.array_equal(vv, [1.8, 1.8])          # DC: two points
    assert np.array_equal(ti, [0, 1e-9, 2 * 10.0 ** -9]) and np.array_equal(vi, [1e-3, 1e-3, 2e-3])  # point added at t = 0
    assert list(N.kind) == [bg.SRC_DC, bg.SRC_PWL]
    assert list(N.ports) == [0, 2]            # ground and unknown names are skipped
    assert (N.node_row("b"), N.node_row("0"), N.node_row("gnd"), N.node_row("zz")) == (1, -1, -1, -2)


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libref_parser.so not built (needs /root/reference)")
@pytest.mark.parametrize("load", ["dc", "pwl", "pulse"])
def test_equals_reference_parser_on_grids(tmp_path, load):
    """power grids in the ibmpg dialect: matrices, node numbering, ports, .tran and source tables bit for bit"""
    text = synth.pg2d_netlist(23, 17, seed=synth.SEED_BASE + 7, pad_pitch=8, nsteps=40, load=load, ports=((0, 0), (5, 3), (22, 16)))
    N, R = assert_matches_reference(text, tmp_path)
    if load == "pulse":   # the records the reference hands to gen_PULSEut_kernel
        assert np.array_equal(N.pulse_time, R["pulse_time"]) and np.array_equal(N.pulse_val, R["pulse_val"])
        assert np.array_equal(N.pulse_src, N.info.n_vsources + np.arange(N.info.n_isources))


def mixed_netlist():
    """inductors, floating and grounded sources in both orientations, every scale letter, parallel and repeated
    elements, continuation lines, upper/lower case keywords.  (A PULSE source sits last in its list: the reference cuts
    a PULSE table that a PWL source follows down to the previous PWL's length, src/parser.cpp:2553-2556 -- a defect
    the product does not share.)"""
    rng = np.random.default_rng(5)
    names = [f"n{k}" for k in range(30)]
    lines = ["* mixed"]
    suffix = ["", "k", "K", "m", "M", "meg", "MEG", "u", "n", "p", "f", "g", "t", "e3", "e-2"]
    for k in range(120):
        a, b = rng.choice(31, 2, replace=False)
        na, nb = (names[a] if a < 30 else "0"), (names[b] if b < 30 else "gnd")
        val = f"{rng.uniform(0.5, 9.5):.4f}{suffix[k % len(suffix)]}"
        lines.append(f"{'RrCcLl'[k % 6]}{k} {na} {nb} {val}")
    lines += ["V1 n0 0 1.8", "v2 0 n1 0.9", "V3 n2 n3 PWL(0 0 1n 1 2n 0.5)", "V4 n4 0 PWL(", "+ 1n 1.0", "+ 2n 2.0", "+ )",
              "V5 n5 gnd 0 PULSE(0 1.2 1n 0.1n 0.1n 2n 5n)",
              "I1 n6 0 PWL(0 0 1n 1m 3n 0)", "i2 0 n7 pwl(0.5n 2m 1n 3m)", "I4 n10 0 PWL(0 1m 4n 2m)", "I3 n8 n9 0 PULSE(0 5m 0 1n 1n 3n 10n)",
              ".TRAN 0.1n 20n", ".print tran v(n3) v(n29) v(n0)", ".end"]
    return "\n".join(lines) + "\n"


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libref_parser.so not built (needs /root/reference)")
def test_equals_reference_parser_on_mixed_elements(tmp_path):
    """inductors, floating and grounded sources in both orientations, every scale letter, parallel and repeated
    elements (duplicate stamps add up in file order), continuation lines, upper/lower case keywords"""
    N, R = assert_matches_reference(mixed_netlist(), tmp_path)
    assert np.array_equal(N.pulse_time, [[1e-9, 0.1 * 10.0 ** -9, 0.1 * 10.0 ** -9, 2 * 10.0 ** -9, 5 * 10.0 ** -9], [0, 1e-9, 1e-9, 3 * 10.0 ** -9, 10 * 10.0 ** -9]])
    assert list(N.pulse_src) == [4, 5 + 3]
    assert np.array_equal(R["pulse_time"][0], N.pulse_time[1])   # the reference records current sources only


def test_reduced_formulation_solves_the_same_circuit():
    """eliminating grounded voltage sources: same node voltages as the full MNA system (direct solves), the reduced G
    of the generated grid IS synth.pg2d's matrix, and the transient source vector follows the tables"""
    nx, ny = 40, 35
    text = synth.pg2d_netlist(nx, ny, seed=synth.SEED_BASE + 2, nsteps=20, load="pwl")
    N, R = bg.Netlist(text, text=True), bg.Netlist(text, formulation=bg.NETLIST_REDUCED, text=True)
    assert R.info.n_eliminated == N.info.n_vsources and R.info.size == nx * ny and R.info.size_mna == N.info.size
    for t in (0.0, 7e-11, 1.5e-10):
        x = spl.spsolve(dense(N, "G").tocsc(), dense(N, "B") @ N.u(t))
        xr = spl.spsolve(dense(R, "G").tocsc(), dense(R, "B") @ R.u(t))
        full = R.expand(xr, t)
        nn = N.info.n_nodes
        assert np.abs(full[:nn] - x[:nn]).max() <= 1e-12 * np.abs(x[:nn]).max()
    rp, ci, v, b = synth.pg2d(nx, ny, seed=synth.SEED_BASE + 2)
    A = sp.csr_matrix((v, ci, rp), shape=(nx * ny, nx * ny))
    assert abs(dense(R, "G") - A).max() <= 4e-16 * abs(A).max()
    assert np.abs(dense(R, "B") @ R.u(1.0) - b).max() <= 1e-15     # loads fully ramped up
    assert np.array_equal(dense(R, "C").diagonal(), dense(N, "C").diagonal()[:nx * ny])
    assert list(R.ports) == [0] and R.node_row("_X_n_8_8") == -3
    # a floating source keeps its MNA row; a fixed node coupled by a capacitor to the grid needs a DC source
    M = bg.Netlist("R1 a b 1\nR2 b 0 1\nV1 a b 1\nV2 c 0 2\nR3 c a 1\n.end\n", formulation=bg.NETLIST_REDUCED, text=True)
    assert (M.info.size, M.info.n_eliminated) == (3, 1)
    xm = spl.spsolve(dense(M, "G").tocsc(), dense(M, "B") @ M.u(0))
    F = bg.Netlist("R1 a b 1\nR2 b 0 1\nV1 a b 1\nV2 c 0 2\nR3 c a 1\n.end\n", text=True)
    xf = spl.spsolve(dense(F, "G").tocsc(), dense(F, "B") @ F.u(0))
    assert np.allclose(M.expand(xm)[:3], xf[:3], rtol=0, atol=1e-14)
    with pytest.raises(bg.BgError, match="du/dt"):
        bg.Netlist("V1 a 0 PWL(0 0 1n 1)\nC1 a b 1p\nR1 b 0 1\n.tran 1p 1n\n.end\n", formulation=bg.NETLIST_REDUCED, text=True)
    with pytest.raises(bg.BgError, match="two voltage sources"):
        bg.Netlist("V1 a 0 1\nV2 a 0 2\nR1 a 0 1\n.end\n", formulation=bg.NETLIST_REDUCED, text=True)


REF_CONV_SO = os.path.join(ROOT, "oracle", "_ref", "libref.so")


@pytest.mark.skipif(not (os.path.exists(REF_SO) and os.path.exists(REF_CONV_SO)), reason="oracle/_ref not built (needs /root/reference)")
def test_handover_to_the_solver_equals_the_reference_chain(tmp_path):
    """the reference hands its parser's cs_dl matrices to the GPU solver through
    LDcsc2csrMySpMatrix (src/formatConvert.cpp:300-334: COO, cycle-chasing COO->CSR, per-row bubble sort, values cast
    to float).  The CSR bg_netlist_get_matrix returns, cast to float, is that matrix element for element."""
    L = C.CDLL(REF_CONV_SO)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for k, text in enumerate((synth.pg2d_netlist(14, 11, seed=synth.SEED_BASE + 41, pad_pitch=4, nsteps=8, load="pwl"), mixed_netlist(), random_netlist(77))):
        path = tmp_path / f"h{k}.sp"
        path.write_text(text)
        N, R = bg.Netlist(path), ref_parse(path)
        for w in "GC":
            M = R[w].tocsc()          # the reference parser's own arrays
            if M.nnz == 0:
                continue
            n = M.shape[0]
            cp, ri, x = M.indptr.astype(np.int64), M.indices.astype(np.int64), np.ascontiguousarray(M.data)
            rp, ci, v = np.zeros(n + 1, np.int32), np.zeros(M.nnz, np.int32), np.zeros(M.nnz, np.float32)
            L.ref_LDcsc2csr(C.c_long(n), C.c_long(n), p(cp), p(ri), p(x), p(rp), p(ci), p(v))
            mine = getattr(N, w)
            assert np.array_equal(rp, mine[0]) and np.array_equal(ci, mine[1]) and np.array_equal(v, mine[2].astype(np.float32)), (k, w)


def random_netlist(seed):
    """a random flat netlist inside the dialect both parsers accept (no PWL left open on its first line, no PULSE source
    followed by a PWL one -- the reference's defects listed in DESIGN.md)"""
    rng = np.random.default_rng(seed)
    nn = int(rng.integers(3, 25))
    node = lambda: ("0", "gnd")[int(rng.integers(2))] if rng.uniform() < 0.15 else f"{('n', 'N', 'x_', 'net')[int(rng.integers(4))]}{int(rng.integers(nn))}"
    sfx = ["", "k", "m", "MEG", "u", "n", "p", "f", "G", "T", "e-3", "E2"]
    num = lambda: f"{rng.uniform(0.1, 99.0):.{int(rng.integers(1, 7))}f}{sfx[int(rng.integers(len(sfx)))]}"
    sep = lambda: (" ", "  ", "\t")[int(rng.integers(3))]
    lines = ["* fuzz %d" % seed]
    n_v = int(rng.integers(0, 4))
    n_i = int(rng.integers(0, 5))
    elems = ["R"] * int(rng.integers(3, 40)) + ["C"] * int(rng.integers(0, 15)) + ["L"] * int(rng.integers(0, 4))
    rng.shuffle(elems)
    # voltage sources before current sources, never interleaved: the reference shares its PWL bookkeeping between the
    # two lists and resizes the wrong table otherwise (src/parser.cpp:2367-2370, 2553-2556)
    cut = sorted(int(x) for x in rng.integers(0, len(elems) + 1, size=2))
    elems = elems[:cut[0]] + ["V"] * n_v + elems[cut[0]:cut[1]] + ["I"] * n_i + elems[cut[1]:]
    tstop = float(rng.choice([0.0, 1e-9, 3.5e-9]))
    seen = dict(V=0, I=0)
    kinds = {"V": ["dc"] * n_v, "I": ["pwl"] * n_i}       # current sources: the reference polls stdin for DC ones
    for which, cnt in (("V", n_v), ("I", n_i)):
        for k in range(cnt):
            r = rng.uniform()
            kinds[which][k] = "pwl" if r < 0.5 else ("dc" if which == "V" and r < 0.8 else "pwl")
        if cnt and tstop > 0 and rng.uniform() < 0.5:
            kinds[which][cnt - 1] = "pulse"                # a PULSE source only as the last of its list
    used = []
    for k, e in enumerate(elems):
        a, b = node(), node()
        while b == a or (a in ("0", "gnd") and b in ("0", "gnd")):
            b = node()
        used += [a, b]
        name = f"{e if rng.uniform() < 0.7 else e.lower()}{k}"
        if e in "RCL":
            lines.append(f"{name}{sep()}{a}{sep()}{b}{sep()}{num()}")
            continue
        kind = kinds[e][seen[e]]
        seen[e] += 1
        if kind == "dc":
            lines.append(f"{name} {a} {b} {num()}")
        elif kind == "pulse":
            per = 1e-9
            lines.append(f"{name} {a} {b} 0 PULSE({num()} {num()} {0.1 * per!r} {0.05 * per!r} {0.07 * per!r} {0.3 * per!r} {per!r})")
        else:
            pts = sorted(float(t) for t in rng.uniform(0, 4e-9, size=int(rng.integers(1, 6))))
            if rng.uniform() < 0.4:
                pts[0] = 0.0
            pairs = [f"{t!r} {num()}" for t in pts]
            if rng.uniform() < 0.5:
                lines.append(f"{name} {a} {b} {'PWL' if rng.uniform() < 0.5 else 'pwl'}({' '.join(pairs)})")
            else:
                lines += [f"{name} {a} {b} PWL("] + [f"+ {p}" for p in pairs] + ["+ )"]
    if tstop > 0:
        lines.append(f".tran {tstop / 50!r} {tstop!r}")
    # observed nodes must exist: for an unknown name the reference indexes its node table at -1 (src/element.cpp:163-175)
    lines.append(".print tran " + " ".join(f"v({used[int(rng.integers(len(used)))]})" for _ in range(int(rng.integers(1, 5)))))
    lines.append(".end")
    return "\n".join(lines) + "\n"


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libref_parser.so not built (needs /root/reference)")
def test_fuzz_against_reference_parser(tmp_path):
    """150 random netlists: element mixes, name spellings, separators, scale letters, grounded / floating / reversed
    sources, both PWL spellings -- every one bit-identical with the reference's compiled parser"""
    for seed in range(150):
        text = random_netlist(1000 + seed)
        try:
            assert_matches_reference(text, tmp_path, name=f"fuzz{seed}.sp")
        except AssertionError as e:
            raise AssertionError(f"seed {1000 + seed}:\n{text}\n{e}") from e


def test_mna_system_is_solvable_the_reference_way():
    """full MNA formulation (zero diagonal on the voltage-source rows) + the reference's ILU++ with pivoting + its
    GMRESilu (CPU oracle): the flow of mna_solve_gpu_gmres.cpp from the parser on, checked against a direct solve"""
    sys.path.insert(0, ROOT)
    from oracle import pyoracle as oracle

    oracle.build()
    if oracle.ilupp() is None:
        pytest.skip("oracle/_ref/libilupp.so not built (needs /root/reference)")
    N = bg.Netlist(synth.pg2d_netlist(40, 36, seed=synth.SEED_BASE + 31, pad_pitch=8), text=True)
    n = N.info.size
    rp, ci, v = N.G
    A = dense(N, "G")
    assert int((A.diagonal() == 0).sum()) == N.info.n_vsources == 20
    b = dense(N, "B") @ N.u(0.0)
    Ac = A.tocsc()
    Ac.sort_indices()
    fac = oracle.ilupp_factor(n, Ac.indptr, Ac.indices, Ac.data, threshold=1.2, mem_factor=1.0)
    keys = ("lp", "li", "lv", "up", "ui", "uv", "middle", "lscale", "rscale", "permRow", "permCol")
    pc = oracle.Precond(oracle.PC_SPLIT, n, "f64", **{k: fac[k] for k in keys})
    flag, xo, it, res, _ = oracle.gmres(rp, ci, v, b, np.zeros(n), 32, 10000, 1e-8, pc)
    oracle.ilupp_free(fac)
    xd = spl.spsolve(Ac, b)
    assert flag == 0 and it < 40 and np.linalg.norm(xo - xd) <= 1e-6 * np.linalg.norm(xd)


def test_source_values_follow_tables_and_pulse_records():
    N = bg.Netlist("R1 a 0 1\nV1 a 0 0 PULSE(0.2 1.2 1n 0.5n 0.25n 2n 5n)\nI1 a 0 PWL(1n 1m 3n 3m)\n.tran 0.1n 7n\n.end\n", text=True)
    (tv, vv), (ti, vi) = N.pwl
    assert len(tv) == 12 and tv[-1] == 7 * 10.0 ** -9 and tv[-2] > tv[-1]      # the reference's cut last period
    for t, want in ((0.0, 0.2), (1.25e-9, 0.7), (2e-9, 1.2), (3.625e-9, 0.7), (4.9e-9, 0.2), (6.25e-9, 0.7), (6.9e-9, 1.2)):
        assert abs(N.u(t)[0] - want) < 1e-12, t
    assert np.allclose(N.u(0.0)[1], 1e-3) and np.allclose(N.u(2e-9)[1], 2e-3) and np.allclose(N.u(9e-9)[1], 3e-3)
    assert [len(t) for t, _ in N.transient_pwl] == [1, 3]


def test_include_and_errors(tmp_path):
    (tmp_path / "sub").mkdir()
    (tmp_path / "sub" / "grid.inc").write_text("R2 b 0 2\nR3 b c 3\n* end of include\n")
    (tmp_path / "top.sp").write_text('R1 a b 1\n.include "sub/grid.inc"\nV1 a 0 1\n.end\nR9 zz 0 1\n')
    N = bg.Netlist(tmp_path / "top.sp")
    assert [N.row_name(r) for r in range(3)] == ["a", "b", "c"] and N.info.n_resistors == 3   # nothing after .end
    flat = bg.Netlist("R1 a b 1\nR2 b 0 2\nR3 b c 3\nV1 a 0 1\n.end\n", text=True)
    assert same_matrix(dense(N, "G"), dense(flat, "G").tocsc())
    for text, code, pat in [("R1 a b\n.end\n", bg.ERR_ARG if hasattr(bg, "ERR_ARG") else -1, "line 1"),
                            ("R1 a b 1\nR2 a b -3\n", -1, "line 2"),
                            ("R1 a b 1\nX1 a b cell\n", -6, "sub-circuit"),
                            (".subckt cell p q\nR1 p q 1\n.ends\n", -6, "sub-circuit"),
                            ("I1 a 0 PWL(0 1 2)\n", -1, "pairs"),
                            ("V1 a 0 0 PULSE(0 1 1n 1n 1n)\n", -1, "PULSE"),
                            ('.include "nosuch.inc"\n', -1, "cannot open")]:
        with pytest.raises(bg.BgError, match=pat) as e:
            bg.Netlist(text, text=True, include_dir=str(tmp_path))
        assert e.value.code == code
    with pytest.raises(bg.BgError, match="cannot open"):
        bg.Netlist(tmp_path / "missing.sp")
    empty = bg.Netlist("* nothing\n.end\n", text=True)
    assert empty.info.size == 0 and empty.info.dc == 1
