"""Synthetic potential-energy surfaces used as benchmark / test *inputs*
(SURVEY.md section 8d).  No optimiser arithmetic lives here.

``quadratic_system(b, n)``: the fixed indefinite-quadratic surface of system ``b``

    f(x) = 1/2 (x - x*)^T A (x - x*),   g(x) = A (x - x*)
    A = Q diag(lam) Q^T,  lam_0 = -0.5 (exactly one negative curvature, i.e. an
    order-1 saddle at x*),  lam_i = 0.1 + |N(0,1)|  otherwise,
    x0 = x* + 0.3 N(0,1)^n / sqrt(n),     rng = RandomState(1000 + b)

``quadratic_batch_torch`` draws the same family directly on a torch device for
the large benchmark configurations (1024 x 384^2 and up), where a per-system
host QR would take minutes.

``fcc_cluster`` / ``fcc111_slab``: copper geometries of the EMT configurations of BASELINE.json
(C2: 64-atom clusters, C3: 128-atom slabs), built without ASE.
"""
import numpy as np


def quadratic_system(b, n, conditioning="normal"):
    rng = np.random.RandomState(1000 + b)
    Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    if conditioning == "loguniform":
        lam = np.exp(rng.uniform(np.log(0.05), np.log(50.0), size=n))
    else:
        lam = 0.1 + np.abs(rng.normal(size=n))
    lam[0] = -0.5
    A = (Q * lam[None, :]) @ Q.T
    A = 0.5 * (A + A.T)
    xstar = rng.normal(size=n)
    x0 = xstar + 0.3 * rng.normal(size=n) / np.sqrt(n)
    return A, xstar, x0


def quadratic_batch(batch, n, first=0, conditioning="normal"):
    """Stacked (A[b,n,n], xstar[b,n], x0[b,n]) for systems first..first+batch-1."""
    A = np.empty((batch, n, n))
    xs = np.empty((batch, n))
    x0 = np.empty((batch, n))
    for i in range(batch):
        A[i], xs[i], x0[i] = quadratic_system(first + i, n, conditioning)
    return A, xs, x0


def quadratic_func(A, xstar):
    """Host callable x -> (f, g) for one system (used by the CPU oracle)."""
    def func(x):
        d = x - xstar
        g = A @ d
        return 0.5 * (d @ g), g
    return func


def quadratic_batch_torch(batch, n, device, seed=1000, chunk=256):
    """Same family, generated with torch on ``device`` (input generation only)."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    A = torch.empty((batch, n, n), dtype=torch.float64, device=device)
    xs = torch.empty((batch, n), dtype=torch.float64, device=device)
    x0 = torch.empty((batch, n), dtype=torch.float64, device=device)
    for lo in range(0, batch, chunk):
        hi = min(batch, lo + chunk)
        m = hi - lo
        G = torch.randn((m, n, n), dtype=torch.float64, device=device, generator=gen)
        Q, _ = torch.linalg.qr(G)
        lam = 0.1 + torch.randn((m, n), dtype=torch.float64, device=device,
                                generator=gen).abs()
        lam[:, 0] = -0.5
        Ab = (Q * lam[:, None, :]) @ Q.transpose(1, 2)
        A[lo:hi] = 0.5 * (Ab + Ab.transpose(1, 2))
        xs[lo:hi] = torch.randn((m, n), dtype=torch.float64, device=device, generator=gen)
        x0[lo:hi] = xs[lo:hi] + 0.3 * torch.randn(
            (m, n), dtype=torch.float64, device=device, generator=gen) / n ** 0.5
        del G, Q, Ab
    return A, xs, x0


# --------------------------------------------------------------------------- copper geometries
def fcc_cluster(natoms, a=3.61, seed=0, rattle=0.05):
    """Ball cut from fcc copper + Gaussian rattle (config C2 of BASELINE.json)."""
    m = 6
    pts = []
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    for i in range(-m, m + 1):
        for j in range(-m, m + 1):
            for k in range(-m, m + 1):
                for bvec in basis:
                    pts.append((np.array([i, j, k]) + bvec) * a)
    pts = np.array(pts) - np.array([0.13, 0.07, 0.03]) * a          # generic centre: no ties at the surface
    order = np.argsort((pts ** 2).sum(1), kind="stable")
    pos = pts[order[:natoms]].copy()
    pos -= pos.mean(0)
    rng = np.random.RandomState(seed)
    return pos + rattle * rng.normal(size=pos.shape)


def fcc111_slab(nx, ny, nlayers, a=3.61, vacuum=7.5, seed=None, rattle=0.05):
    """Orthogonal fcc(111) slab: nx x ny surface cells (2 atoms each per layer), ABC stacking,
    periodic in x and y.  Returns (positions, cell, pbc)."""
    d = a / np.sqrt(2.0)
    ax, ay, dz = d, d * np.sqrt(3.0), a / np.sqrt(3.0)
    pos = []
    for l in range(nlayers):
        off = np.array([0.0, (l % 3) * ay / 3.0])
        for i in range(nx):
            for j in range(ny):
                for bx, by in ((0.0, 0.0), (0.5, 0.5)):
                    p = np.array([(i + bx) * ax, (j + by) * ay]) + off
                    pos.append([p[0] % (nx * ax), p[1] % (ny * ay), vacuum + l * dz])
    pos = np.array(pos)
    cell = np.diag([nx * ax, ny * ay, 2 * vacuum + (nlayers - 1) * dz])
    if seed is not None:
        pos = pos + rattle * np.random.RandomState(seed).normal(size=pos.shape)
    return pos, cell, (True, True, False)


def fcc111_with_adatom(size=(5, 5, 6), a=3.61, vacuum=7.5, height=2.0):
    """The README example of the reference (README.md:19-20; BASELINE.json config C1) without ASE:
    ase.build.fcc111('Cu', size, vacuum) -- hexagonal surface cell, ABC stacking, the top layer an
    fcc continuation of the ones below -- plus one Cu adatom `height` above a bridge site.
    Returns (positions [nx*ny*nl + 1, 3], cell, pbc)."""
    nx, ny, nl = size
    d = a / np.sqrt(2.0)
    dz = a / np.sqrt(3.0)
    a1 = np.array([d, 0.0, 0.0])
    a2 = np.array([0.5 * d, 0.5 * np.sqrt(3.0) * d, 0.0])
    shift = (a1 + a2) / 3.0
    pos = []
    for l in range(nl):
        off = ((nl - 1 - l) % 3) * shift                # top layer at offset 0, going down A, B, C
        for j in range(ny):
            for i in range(nx):
                p = i * a1 + j * a2 + off
                pos.append([p[0], p[1], vacuum + l * dz])
    top = vacuum + (nl - 1) * dz
    bridge = 0.5 * a1                                    # midway between two neighbouring top-layer atoms
    pos.append([bridge[0], bridge[1], top + height])
    cell = np.array([nx * a1, ny * a2, [0.0, 0.0, 2 * vacuum + (nl - 1) * dz]])
    return np.array(pos), cell, (True, True, False)
