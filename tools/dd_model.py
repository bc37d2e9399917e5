"""Executable model of the round-2 spatial decomposition -- numpy, CPU only.

Not product code: it pins down, and lets tests/test_dd_model_cpu.py verify, the three rules the CUDA implementation will
follow so that P ranks reproduce the single-rank forces BIT FOR BIT:

 1. home rank   = the brick (px x py x pz grid over the fractional coordinates of the periodic cell) that holds the atom;
 2. halo        = every atom that is not home but lies within the padded cutoff of the brick (periodic distance to the
                  brick, per axis): if |xi - xj| < r then j is within r of i's brick, so BOTH home ranks of a cross pair
                  hold both atoms;
 3. pair owner  = the home rank if both atoms share it, otherwise the home rank of the atom with the smaller
                  (brick index, atom index) key -- computed from data both candidates hold, no communication;
    the owner evaluates the pair once (Newton's third law) and accumulates both forces in 2^32 fixed point; halo forces
    travel back to their home ranks as int64 and are ADDED there, which is exact and order independent.
"""
import numpy as np

SCALE = 4294967296.0


def brick_of(frac, grid):
    """home brick (linear index) of fractional coordinates in [0,1)^3"""
    g = np.asarray(grid)
    c = np.minimum((frac*g).astype(int), g - 1)
    return (c[:, 0]*g[1] + c[:, 1])*g[2] + c[:, 2]


def brick_bounds(b, grid):
    g = np.asarray(grid)
    c = np.array([b // (g[1]*g[2]), (b // g[2]) % g[1], b % g[2]])
    return c/g, (c + 1)/g


def periodic_gap(frac, lo, hi):
    """per-axis periodic distance (in fractional units) from points to the interval [lo, hi)"""
    d = np.zeros_like(frac)
    for k in range(3):
        x = frac[:, k]
        inside = (x >= lo[k]) & (x < hi[k])
        dl = np.abs(x - lo[k]); dl = np.minimum(dl, 1 - dl)
        dh = np.abs(x - hi[k]); dh = np.minimum(dh, 1 - dh)
        d[:, k] = np.where(inside, 0.0, np.minimum(dl, dh))
    return d


def local_sets(pos, box, grid, rank, reach):
    """(home, halo) atom indices of `rank` for an orthorhombic box; reach = cutoff + padding"""
    L = np.diag(box)
    frac = (pos/L) % 1.0
    home_rank = brick_of(frac, grid)
    lo, hi = brick_bounds(rank, grid)
    gap = periodic_gap(frac, lo, hi)*L
    near = (gap**2).sum(axis=1) < reach*reach
    home = np.where(home_rank == rank)[0]
    halo = np.where(near & (home_rank != rank))[0]
    return home, halo, home_rank


def pair_force(d, r2, qq, sig, eps):
    """Coulomb (plain, truncated) + Lennard-Jones: dE/dr over r, i.e. force on j = -d * this ... returns F_j = f*d"""
    inv2 = 1.0/r2
    s6 = (sig*sig*inv2)**3
    return (qq*np.sqrt(inv2)*inv2 + eps*(12*s6*s6 - 6*s6)*inv2)[:, None]*d


def forces_of_rank(pos, box, charges, sigmas, epsilons, cutoff, grid, rank, reach):
    """int64 fixed-point force array (global atom numbering) holding what `rank` computes: every pair it owns, once"""
    L = np.diag(box)
    home, halo, home_rank = local_sets(pos, box, grid, rank, reach)
    loc = np.concatenate([home, halo])
    out = np.zeros((len(pos), 3), dtype=np.int64)
    if len(loc) < 2:
        return out
    ii, jj = np.triu_indices(len(loc), 1)
    a, b = loc[ii], loc[jj]
    ra, rb = home_rank[a], home_rank[b]
    key_a_smaller = (ra < rb) | ((ra == rb) & (a < b))
    owner = np.where(ra == rb, ra, np.where(key_a_smaller, ra, rb))
    mine = owner == rank
    a, b = a[mine], b[mine]
    d = pos[b] - pos[a]
    d -= L*np.rint(d/L)
    r2 = (d*d).sum(axis=1)
    ok = r2 < cutoff*cutoff
    a, b, d, r2 = a[ok], b[ok], d[ok], r2[ok]
    f = pair_force(d, r2, 138.935458*charges[a]*charges[b], 0.5*(sigmas[a] + sigmas[b]), 4*np.sqrt(epsilons[a]*epsilons[b]))
    fx = np.rint(f*SCALE).astype(np.int64)                  # per-PAIR rounding: the sum no longer depends on who adds what
    np.add.at(out, b, fx)
    np.add.at(out, a, -fx)
    return out


def model_system(n=240, L=2.4, seed=5):
    rng = np.random.default_rng(seed)
    m = int(round(n**(1/3) + 0.5))
    grid = np.stack(np.meshgrid(*[np.arange(m)]*3, indexing="ij"), -1).reshape(-1, 3)[:n]
    pos = (grid + 0.5)*(L/m) + rng.uniform(-0.08, 0.08, size=(n, 3))
    q = np.where(np.arange(n) % 2 == 0, 0.4, -0.4)
    return pos % L, np.diag([L, L, L]), q, np.full(n, 0.25), np.full(n, 0.5)
