"""Generate the committed golden fixtures of the front-end rows (netlist -> MNA, ILU(k)) under tests/golden/.

    python tests/golden/make_frontend_golden.py        (build container only: needs oracle/_ref built from /root/reference)

* netlist_*.npz : the output of the REFERENCE's own parser (src/parser.cpp parser_sub + stampB/stampC/stampG, compiled
  unchanged into oracle/_ref/libref_parser.so) on the netlist text stored in the same file: column-compressed G, C, B,
  node names in row order, ports, .tran values, every source table.
* iluk_*.npz    : the ILU(k) of the ITSOL code the reference vendors (src/iluk.cpp, oracle/_ref/libref_iluk.so) on a
  seeded matrix stored in the same file, levels 0-3: per-row counts, columns, float values, reciprocal diagonal.
tests/test_netlist.py and tests/test_host_logic.py compare the product with these fixtures wherever they run, and with the
compiled reference itself where oracle/_ref exists.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "gpu-gmres_b200", "python"))
from b200gmres import synth  # noqa: E402
import test_netlist as TN  # noqa: E402
import test_host_logic as TH  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def netlist_cases():
    yield "netlist_grid_pwl", synth.pg2d_netlist(9, 7, seed=synth.SEED_BASE + 21, pad_pitch=4, nsteps=12, load="pwl", ports=((0, 0), (4, 3), (8, 6)))
    yield "netlist_grid_pulse", synth.pg2d_netlist(8, 8, seed=synth.SEED_BASE + 22, pad_pitch=4, nsteps=40, load="pulse", ports=((1, 1),))
    yield "netlist_mixed", TN.mixed_netlist()


def main():
    import tempfile

    for name, text in netlist_cases():
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "case.sp")
            with open(path, "w") as f:
                f.write(text)
            R = TN.ref_parse(path)
        out = dict(text=np.array(text), counts=np.array([R["n_nodes"], R["n_l"], R["n_vs"], R["n_is"]], np.int64),
                   tran=np.array([R["tstep"], R["tstop"]]), ports=R["ports"], names=np.array(R["names"]),
                   pulse_time=R["pulse_time"], pulse_val=R["pulse_val"],
                   src_ptr=np.cumsum([0] + [len(t) for t, _ in R["sources"]]).astype(np.int64),
                   src_t=np.concatenate([t for t, _ in R["sources"]]) if R["sources"] else np.zeros(0),
                   src_v=np.concatenate([v for _, v in R["sources"]]) if R["sources"] else np.zeros(0))
        for w in "GCB":
            M = R[w].tocsc()
            out[w + "_p"], out[w + "_i"], out[w + "_x"], out[w + "_shape"] = M.indptr.astype(np.int64), M.indices.astype(np.int64), M.data, np.array(M.shape, np.int64)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, R["size"], "unknowns")
    for name, (rp, ci, v, _) in {"iluk_pg2d_13x11": synth.pg2d(13, 11, seed=synth.SEED_BASE + 23), "iluk_th3d_6x5x4": synth.th3d(6, 5, 4, seed=synth.SEED_BASE + 24)}.items():
        out = dict(rowptr=np.asarray(rp, np.int32), col=np.asarray(ci, np.int32), val=np.asarray(v, np.float64))
        for level in range(4):
            nl, lc, lv, nu, uc, uv, dinv = TH.itsol_iluk(rp, ci, v, level)
            for k, a in (("nl", nl), ("lc", lc), ("lv", lv), ("nu", nu), ("uc", uc), ("uv", uv), ("dinv", dinv)):
                out[f"k{level}_{k}"] = a
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, len(rp) - 1, "rows")


if __name__ == "__main__":
    main()
