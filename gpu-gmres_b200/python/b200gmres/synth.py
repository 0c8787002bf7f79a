"""Seeded synthetic workloads for the GMRES(m)+ILU hot path.

The reference ships no matrices, so tests and bench.py run on these generators:

* ``pg2d``  -- 2-D power-grid MNA (5-point), VDD pads as Norton equivalents so that A is an
  irreducibly diagonally dominant M-matrix.  Row index = iy*nx + ix, the natural
  order the reference's drivers produce (netlist node order).
* ``th3d``  -- 3-D IC thermal grid with micro-channel convection (7-point, non-symmetric),
  assembled as G + C/h like src_thermal/parserGCB.m:5.

Values are generated in float64 with numpy's PCG64 and cast once by the caller.
CSR arrays are int32 with columns ascending in every row (what formatConvert.cpp:112-163 yields).
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 0x5EED0000


def _coo_to_csr(n, rows, cols, vals):
    """Sort COO triplets (no duplicates) into CSR with ascending columns."""
    key = rows.astype(np.int64) * n + cols.astype(np.int64)
    order = np.argsort(key, kind="stable")
    rows, cols, vals = rows[order], cols[order], vals[order]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    rowptr[1:] = np.bincount(rows, minlength=n)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    return rowptr, cols.astype(np.int32), vals.astype(np.float64)


def pg2d(nx: int, ny: int, seed: int = SEED_BASE + 1, pad_pitch: int = 16, c_over_h: float = 0.0):
    """2-D power grid.  Returns (rowptr, col, val, b) with val/b float64.

    g_h = 1.0*(1+0.1u), g_v = 0.5*(1+0.1u), u~U(-1,1); pads (10 S to 1.8 V) every ``pad_pitch``
    nodes in both directions; current loads U(0,1e-3) A on a random 20 % of the nodes.
    ``c_over_h`` > 0 adds the backward-Euler term C/h with C~U(1,2)*c_over_h on the diagonal.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    n = nx * ny
    gh = 1.0 * (1.0 + 0.1 * rng.uniform(-1, 1, size=(ny, nx - 1)))  # (iy, ix)-(iy, ix+1)
    gv = 0.5 * (1.0 + 0.1 * rng.uniform(-1, 1, size=(ny - 1, nx)))  # (iy, ix)-(iy+1, ix)
    diag = np.zeros((ny, nx))
    diag[:, :-1] += gh
    diag[:, 1:] += gh
    diag[:-1, :] += gv
    diag[1:, :] += gv
    b = np.zeros((ny, nx))
    off = pad_pitch // 2
    pad = np.zeros((ny, nx), dtype=bool)
    pad[off::pad_pitch, off::pad_pitch] = True
    diag[pad] += 10.0
    b[pad] += 10.0 * 1.8
    load = rng.uniform(0, 1, size=(ny, nx)) < 0.2
    b[load] -= rng.uniform(0, 1e-3, size=int(load.sum()))
    if c_over_h > 0:
        diag += c_over_h * rng.uniform(1, 2, size=(ny, nx))
    idx = np.arange(n, dtype=np.int64).reshape(ny, nx)
    rows = [idx.ravel(), idx[:, :-1].ravel(), idx[:, 1:].ravel(), idx[:-1, :].ravel(), idx[1:, :].ravel()]
    cols = [idx.ravel(), idx[:, 1:].ravel(), idx[:, :-1].ravel(), idx[1:, :].ravel(), idx[:-1, :].ravel()]
    vals = [diag.ravel(), -gh.ravel(), -gh.ravel(), -gv.ravel(), -gv.ravel()]
    rowptr, col, val = _coo_to_csr(n, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))
    return rowptr, col, val, b.ravel().copy()


def pg2d_netlist(nx: int, ny: int, seed: int = SEED_BASE + 1, pad_pitch: int = 16, cap: float = 1e-12, tstep: float = 1e-11,
                 nsteps: int = 0, load: str = "dc", ports=((0, 0),)):
    """The pg2d() grid as an ibmpg-style SPICE netlist (text), for the netlist front end (bg_netlist_*).

    Same random stream and therefore the same conductances, pads and loads as pg2d(): node n_{ix}_{iy} is grid node
    iy*nx+ix and appears first on its capacitor line, so the parser's first-appearance numbering IS pg2d's row order;
    a pad is an ideal 1.8 V source behind a 0.1 ohm package resistor (the Norton pad of pg2d after elimination);
    a load is a current source node -> ground, DC or (load="pwl") a ramp 0 -> I over 10 steps, or (load="pulse") a
    periodic trapezoid.  Capacitors to ground are U(1,2)*cap like pg2d(c_over_h = cap/tstep).  nsteps > 0 adds
    ".tran tstep nsteps*tstep"; `ports` become ".print tran v(...)".
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    gh = 1.0 * (1.0 + 0.1 * rng.uniform(-1, 1, size=(ny, nx - 1)))
    gv = 0.5 * (1.0 + 0.1 * rng.uniform(-1, 1, size=(ny - 1, nx)))
    off = pad_pitch // 2
    pad = np.zeros((ny, nx), dtype=bool)
    pad[off::pad_pitch, off::pad_pitch] = True
    is_load = rng.uniform(0, 1, size=(ny, nx)) < 0.2
    amps = np.zeros((ny, nx))
    amps[is_load] = rng.uniform(0, 1e-3, size=int(is_load.sum()))
    cval = cap * rng.uniform(1, 2, size=(ny, nx))
    name = lambda ix, iy: f"n_{ix}_{iy}"
    out = [f"* pg2d {nx} x {ny}, seed {seed:#x}: synthetic power grid in the ibmpg netlist dialect"]
    for iy in range(ny):
        for ix in range(nx):
            out.append(f"C{iy * nx + ix} {name(ix, iy)} 0 {float(cval[iy, ix])!r}")
    k = 0
    for iy in range(ny):
        for ix in range(nx - 1):
            out.append(f"Rh{k} {name(ix, iy)} {name(ix + 1, iy)} {float(1.0 / gh[iy, ix])!r}"); k += 1
    k = 0
    for iy in range(ny - 1):
        for ix in range(nx):
            out.append(f"Rv{k} {name(ix, iy)} {name(ix, iy + 1)} {float(1.0 / gv[iy, ix])!r}"); k += 1
    k = 0
    for iy, ix in zip(*np.nonzero(pad)):
        out.append(f"Rp{k} {name(ix, iy)} _X_{name(ix, iy)} 0.1")
        out.append(f"Vp{k} _X_{name(ix, iy)} 0 1.8"); k += 1
    k = 0
    tstop = nsteps * tstep
    for iy, ix in zip(*np.nonzero(is_load)):
        a = float(amps[iy, ix])
        if load == "pwl":
            if k % 2:  # both spellings the reference reads: one line, or "PWL(" followed by one pair per "+" line
                out.append(f"Il{k} {name(ix, iy)} 0 PWL(0 0 {10 * tstep!r} {a!r} {max(tstop, 11 * tstep)!r} {a!r})")
            else:
                out += [f"Il{k} {name(ix, iy)} 0 PWL(", f"+ 0 0", f"+ {10 * tstep!r} {a!r}", f"+ {max(tstop, 11 * tstep)!r} {a!r}", "+ )"]
        elif load == "pulse":
            out.append(f"Il{k} {name(ix, iy)} 0 0 PULSE(0 {a!r} {2 * tstep!r} {3 * tstep!r} {3 * tstep!r} {4 * tstep!r} {16 * tstep!r})")
        else:
            out.append(f"Il{k} {name(ix, iy)} 0 {a!r}")
        k += 1
    if nsteps > 0:
        out.append(f".tran {tstep!r} {tstop!r}")
    if ports:
        out.append(".print tran " + " ".join(f"v({name(ix, iy)})" for ix, iy in ports))
    out.append(".end")
    return "\n".join(out) + "\n"


def th3d(nx: int, ny: int, nz: int, seed: int = SEED_BASE + 3, c_over_h: float = 0.05):
    """3-D IC thermal grid, 7-point, non-symmetric on the micro-channel layers (z % 4 == 1).

    Layer conductivities repeat {1.3, 0.04, 1.3, 4.0}; channel layers add -/+ a_cv = 0.3*g_x
    to the x-neighbours (central-difference convection, docs/tvlsi_3d_thermal.pdf eq. (3));
    top surface conductance 0.5; diagonal keeps row sums >= 0 plus C/h.  Power map U(0,1) on
    the two layers z = 0 and z = 2 (mod nz).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    n = nx * ny * nz
    kz = np.array([1.3, 0.04, 1.3, 4.0])[np.arange(nz) % 4]
    idx = np.arange(n, dtype=np.int64).reshape(nz, ny, nx)
    k3 = np.broadcast_to(kz[:, None, None], (nz, ny, nx))
    jit = 1.0 + 0.05 * rng.uniform(-1, 1, size=(nz, ny, nx))
    gx = (k3 * jit)[:, :, :-1]
    gy = (k3 * jit)[:, :-1, :]
    gz = (2.0 * kz[:-1] * kz[1:] / (kz[:-1] + kz[1:]))[:, None, None] * jit[:-1]
    diag = np.zeros((nz, ny, nx))
    diag[:, :, :-1] += gx; diag[:, :, 1:] += gx
    diag[:, :-1, :] += gy; diag[:, 1:, :] += gy
    diag[:-1] += gz; diag[1:] += gz
    diag[-1] += 0.5
    diag += c_over_h
    chan = (np.arange(nz) % 4 == 1)[:, None, None]
    acv = 0.3 * gx * chan  # zero off the channel layers
    rows = [idx.ravel(), idx[:, :, :-1].ravel(), idx[:, :, 1:].ravel(), idx[:, :-1, :].ravel(),
            idx[:, 1:, :].ravel(), idx[:-1].ravel(), idx[1:].ravel()]
    cols = [idx.ravel(), idx[:, :, 1:].ravel(), idx[:, :, :-1].ravel(), idx[:, 1:, :].ravel(),
            idx[:, :-1, :].ravel(), idx[1:].ravel(), idx[:-1].ravel()]
    vals = [diag.ravel(), (-gx + acv).ravel(), (-gx - acv).ravel(), -gy.ravel(), -gy.ravel(),
            -gz.ravel(), -gz.ravel()]
    rowptr, col, val = _coo_to_csr(n, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))
    b = np.zeros((nz, ny, nx))
    for z in (0, 2 % nz):
        b[z] = rng.uniform(0, 1, size=(ny, nx))
    return rowptr, col, val, b.ravel().copy()


def random_sparse(n: int, avg_nnz: int, seed: int, diag_boost: float = 2.0):
    """Irregular diagonally dominant test matrix (ragged rows, a few long ones)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = rng.poisson(max(avg_nnz - 1, 0), size=n)
    lens[rng.integers(0, n, size=max(1, n // 50))] += 10 * avg_nnz  # long rows
    lens = np.minimum(lens, n - 1)
    rows = np.repeat(np.arange(n), lens)
    cols = rng.integers(0, n, size=rows.size)
    keep = rows != cols
    rows, cols = rows[keep], cols[keep]
    key = np.unique(rows.astype(np.int64) * n + cols)
    rows, cols = (key // n).astype(np.int64), (key % n).astype(np.int64)
    vals = rng.uniform(-1, 1, size=rows.size)
    rowsum = np.bincount(rows, weights=np.abs(vals), minlength=n)
    d = diag_boost * (rowsum + 1.0)
    rowptr, col, val = _coo_to_csr(n, np.concatenate([rows, np.arange(n)]),
                                   np.concatenate([cols, np.arange(n)]), np.concatenate([vals, d]))
    return rowptr, col, val, rng.uniform(-1, 1, size=n)
