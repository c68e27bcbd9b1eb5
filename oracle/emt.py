"""CPU restatement (numpy) of an EMT-form copper potential -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

PARITY UNPINNED against ASE: the configurations of BASELINE.json name
``ase.calculators.emt.EMT`` (reference call sites README.md:14,29,
tests/test_peswrapper.py:6,23), but ASE is a third-party dependency (pyproject.toml:28,
``ase>=3.18.0``) that is neither vendored in the reference tree nor installed here, and
none of the reference's tests pins an EMT energy or force.  This
file therefore restates the published functional form (Jacobsen, Stoltze, Norskov,
Surf. Sci. 366 (1996) 394) with the copper parameters of that paper and is the single
definition shared by the CPU oracle loop and the CUDA kernel (sella_b200/csrc/emt.cu),
so optimiser parity is checked on identical surfaces.  "EMT" below always means this
EMT-form potential, not numerical equality with ASE.

    E = sum_i [ E0 ((1 + lam ds_i) exp(-lam ds_i) - 1) + 6 V0 exp(-kappa ds_i) ]
        - 1/2 V0/gamma2 sum_{i != j} w(r_ij) exp(-kappa (r_ij/beta - s0))
    ds_i = -ln(sigma1_i / 12) / (beta eta2)
    sigma1_i = 1/gamma1 sum_{j != i} w(r_ij) exp(-eta2 (r_ij - beta s0))
    w(r) = 1 / (1 + exp(acut (r - rc)))          (smooth cutoff between the 3rd and 4th fcc shell)
"""
import numpy as np

BOHR = 0.52917721
BETA = 1.809                               # (16 pi / 3)^(1/3) / sqrt(2)
# copper: E0 [eV], s0 [A], V0 [eV], eta2 [1/A], kappa [1/A], lambda [1/A]
CU = dict(E0=-3.51, s0=2.67 * BOHR, V0=2.476, eta2=1.652 / BOHR, kappa=2.74 / BOHR, lam=1.906 / BOHR)


def derived(par=CU):
    """rc, acut, gamma1, gamma2, list cutoff (all from the six material constants)."""
    s0, eta2, kappa = par["s0"], par["eta2"], par["kappa"]
    rc = BETA * s0 * 0.5 * (np.sqrt(3.0) + 2.0)
    rr = 4.0 * rc / (np.sqrt(3.0) + 2.0)
    acut = np.log(9999.0) / (rr - rc)
    g1 = g2 = 0.0
    for i, nn in enumerate((12, 6, 24)):
        r = s0 * BETA * np.sqrt(i + 1.0)
        x = nn / (12.0 * (1.0 + np.exp(acut * (r - rc))))
        g1 += x * np.exp(-eta2 * (r - BETA * s0))
        g2 += x * np.exp(-kappa / BETA * (r - BETA * s0))
    return dict(rc=rc, acut=acut, gamma1=g1, gamma2=g2, rlist=rc + 0.5)


def image_ranges(cell, pbc, rlist):
    """How many periodic images are needed along each lattice vector."""
    cell = np.asarray(cell, float).reshape(3, 3)
    out = [0, 0, 0]
    vol = abs(np.linalg.det(cell))
    for d in range(3):
        if not pbc[d]:
            continue
        a, b = cell[(d + 1) % 3], cell[(d + 2) % 3]
        height = vol / np.linalg.norm(np.cross(a, b))
        out[d] = int(np.ceil(rlist / height))
    return out


def emt(x, cell=None, pbc=(False, False, False), par=CU):
    """Energy and gradient dE/dx (flat, 3N) of one configuration."""
    d = derived(par)
    E0, s0, V0, eta2, kappa, lam = (par[k] for k in ("E0", "s0", "V0", "eta2", "kappa", "lam"))
    rc, acut, g1, g2, rlist = d["rc"], d["acut"], d["gamma1"], d["gamma2"], d["rlist"]
    pos = np.asarray(x, float).reshape(-1, 3)
    N = len(pos)
    if cell is None:
        cell = np.zeros((3, 3))
        pbc = (False, False, False)
    cell = np.asarray(cell, float).reshape(3, 3)
    nimg = image_ranges(cell, pbc, rlist) if any(pbc) else [0, 0, 0]
    shifts = [i * cell[0] + j * cell[1] + k * cell[2]
              for i in range(-nimg[0], nimg[0] + 1)
              for j in range(-nimg[1], nimg[1] + 1)
              for k in range(-nimg[2], nimg[2] + 1)]
    sigma1 = np.zeros(N)
    epair = 0.0
    pairs = []                                     # (i, j, unit vector i<-j, r) of every neighbour image
    for sh in shifts:
        dvec = pos[:, None, :] - (pos[None, :, :] + sh[None, None, :])      # r_i - (r_j + shift)
        r = np.sqrt((dvec ** 2).sum(-1))
        mask = (r < rlist) & (r > 1e-9)
        ii, jj = np.nonzero(mask)
        rr = r[ii, jj]
        w = 1.0 / (1.0 + np.exp(acut * (rr - rc)))
        np.add.at(sigma1, ii, w * np.exp(-eta2 * (rr - BETA * s0)) / g1)
        epair += (-0.5 * V0 / g2 * w * np.exp(-kappa * (rr / BETA - s0))).sum()
        pairs.append((ii, jj, dvec[ii, jj] / rr[:, None], rr, w))
    ds = -np.log(sigma1 / 12.0) / (BETA * eta2)
    xl = lam * ds
    ecoh = E0 * ((1.0 + xl) * np.exp(-xl) - 1.0) + 6.0 * V0 * np.exp(-kappa * ds)
    energy = ecoh.sum() + epair
    dF = (E0 * lam * xl * np.exp(-xl) + 6.0 * V0 * kappa * np.exp(-kappa * ds)) / (BETA * eta2 * sigma1)
    grad = np.zeros_like(pos)
    for ii, jj, u, rr, w in pairs:
        dw = -acut * w * (1.0 - w)                  # w'(r)
        e1 = np.exp(-eta2 * (rr - BETA * s0)) / g1
        e2 = np.exp(-kappa * (rr / BETA - s0)) / g2
        drho = dw * e1 - eta2 * w * e1
        dphi = -V0 * (dw * e2 - kappa / BETA * w * e2)
        # every ordered pair (i, j) appears once in the list: atom i feels F'_i rho' + 1/2 phi' from
        # its own sums and F'_j rho' + 1/2 phi' from j's
        coef = (dF[ii] + dF[jj]) * drho + dphi
        np.add.at(grad, ii, coef[:, None] * u)
    return energy, grad.ravel()


def emt_func(cell=None, pbc=(False, False, False), par=CU):
    return lambda x: emt(x, cell, pbc, par)
