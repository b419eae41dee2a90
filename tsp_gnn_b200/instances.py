"""Instance synthesis, the ``.graph`` reader/writer and the batch builder.

Host-side mirror of the reference's data plane for the hot path's input format:
  * ``create_graph``        dataset.py:52-116  (Concorde replaced by a deterministic
                            nearest-neighbour + 2-opt tour: the exact solver is not
                            available offline; the tour only sets the scalar C)
  * ``write_graph/read_graph``  dataset.py:145-187 / instance_loader.py:95-127
  * ``InstanceLoader``      instance_loader.py:7-93, same method names and the same
                            6-tuple ``(EV, W, C, route_exists, n_vertices, n_edges)``

The one deliberate difference: ``EV`` is an :class:`Incidence` (edge_src/edge_dst
columns of the two non-zeros of every row) instead of a dense ``[sumE, sumV]`` array.
``Incidence.toarray()`` gives the dense matrix the reference builds
(instance_loader.py:45,63-66); ``Session.run`` accepts either.
"""
import os
import random
import numpy as np


class Incidence(object):
    """Edge-vertex incidence matrix with exactly two 1.0 entries per row.

    Row e has ones in columns ``src[e] < dst[e]`` (global vertex ids), the layout
    produced by instance_loader.py:52-67.
    """

    def __init__(self, src, dst, n_cols):
        self.src = np.ascontiguousarray(src, dtype=np.int32)
        self.dst = np.ascontiguousarray(dst, dtype=np.int32)
        self.shape = (int(self.src.shape[0]), int(n_cols))

    def toarray(self, dtype=np.float64):
        EV = np.zeros(self.shape, dtype=dtype)
        r = np.arange(self.shape[0])
        EV[r, self.src] = 1
        EV[r, self.dst] = 1
        return EV

    @staticmethod
    def from_dense(EV):
        """Dense 0/1 ``[sumE,sumV]`` -> Incidence; validates two non-zeros per row."""
        EV = np.asarray(EV)
        if EV.ndim != 2:
            raise ValueError("EV must be a matrix")
        r, c = np.nonzero(EV)
        if len(r) != 2 * EV.shape[0] or np.any(r[0::2] != r[1::2]) or np.any(r[0::2] != np.arange(EV.shape[0])):
            raise ValueError("EV must have exactly two non-zeros in every row (instance_loader.py:63-66)")
        return Incidence(c[0::2], c[1::2], EV.shape[1])


# ----------------------------------------------------------------------------
# tours (stand-in for Concorde, dataset.py:9-50)
# ----------------------------------------------------------------------------
def _tour_cost(Mw, route):
    r = np.asarray(route)
    return float(Mw[r, np.roll(r, -1)].sum())


def heuristic_tour(Mw, two_opt_sweeps=3):
    """Deterministic nearest-neighbour tour improved by a few vectorised 2-opt sweeps."""
    n = Mw.shape[0]
    unvisited = np.ones(n, dtype=bool)
    route = [0]
    unvisited[0] = False
    for _ in range(n - 1):
        dcur = np.where(unvisited, Mw[route[-1]], np.inf)
        nxt = int(np.argmin(dcur))
        route.append(nxt)
        unvisited[nxt] = False
    r = np.array(route)
    for _ in range(two_opt_sweeps):
        improved = False
        for i in range(n - 2):
            a, b = r[i], r[i + 1]
            js = np.arange(i + 2, n if i > 0 else n - 1)
            if len(js) == 0:
                continue
            c, dd = r[js], r[(js + 1) % n]
            delta = Mw[a, c] + Mw[b, dd] - Mw[a, b] - Mw[c, dd]
            k = int(np.argmin(delta))
            if delta[k] < -1e-12:
                j = int(js[k])
                r[i + 1:j + 1] = r[i + 1:j + 1][::-1]
                improved = True
        if not improved:
            break
    return [int(v) for v in r]


# ----------------------------------------------------------------------------
# dataset.py:52-116
# ----------------------------------------------------------------------------
def create_graph(n, connectivity=1.0, distances="euc_2D", rng=None, two_opt_sweeps=3):
    """Returns (Ma upper-triangular, Mw, route, nodes) like dataset.create_graph.

    ``rng`` is a ``np.random.RandomState`` (the reference uses the global numpy RNG).
    For connectivity < 1 the planted Hamiltonian cycle (dataset.py:103-107) is returned
    as the route.
    """
    rng = rng if rng is not None else np.random
    Ma = np.zeros((n, n))
    if connectivity >= 1:
        Ma[np.triu_indices(n, 1)] = 1
        Ma = Ma + Ma.T
    else:
        iu = np.triu_indices(n, 1)
        Ma[iu] = (rng.rand(len(iu[0])) < connectivity).astype(float)
        Ma = Ma + Ma.T
    nodes = None
    if distances == "euc_2D":
        nodes = rng.rand(n, 2)                                           # dataset.py:70
        diff = nodes[:, None, :] - nodes[None, :, :]
        Mw = np.sqrt((diff ** 2).sum(-1))                                # dataset.py:73
    elif distances == "random":
        Mw = np.zeros((n, n))
        iu = np.triu_indices(n, 1)
        Mw[iu] = rng.rand(len(iu[0]))
        Mw = Mw + Mw.T
    else:
        raise ValueError("distances must be euc_2D or random")
    if connectivity >= 1:
        route = heuristic_tour(Mw, two_opt_sweeps)
    else:
        perm = [int(v) for v in rng.permutation(n)]
        for i, j in zip(perm, perm[1:] + perm[:1]):
            Ma[i, j] = Ma[j, i] = 1
        route = perm
    return np.triu(Ma), Mw, route, nodes


def write_graph(Ma, Mw, filepath, route=None):
    """dataset.py:145-187 (float weights; the int_weights branch only feeds Concorde)."""
    n = Ma.shape[0]
    with open(filepath, "w") as out:
        out.write("TYPE : TSP\n")
        out.write("DIMENSION: {n}\n".format(n=n))
        out.write("EDGE_DATA_FORMAT: EDGE_LIST\n")
        out.write("EDGE_WEIGHT_TYPE: EXPLICIT\n")
        out.write("EDGE_WEIGHT_FORMAT: FULL_MATRIX \n")
        out.write("EDGE_DATA_SECTION:\n")
        ii, jj = np.nonzero(Ma)
        for i, j in zip(ii, jj):
            out.write("{} {}\n".format(i, j))
        out.write("-1\n")
        out.write("EDGE_WEIGHT_SECTION:\n")
        for i in range(n):
            out.write(" ".join(repr(float(Mw[i, j])) if Ma[i, j] == 1 else "0" for j in range(n)))
            out.write(" \n")
        if route is not None:
            out.write("TOUR_SECTION:\n")
            out.write("{}\n".format(" ".join(str(x) for x in route)))
        out.write("EOF\n")


def read_graph(filepath):
    """instance_loader.py:95-127."""
    with open(filepath, "r") as f:
        line = ""
        while "DIMENSION" not in line:
            line = f.readline()
        n = int(line.split()[1])
        Ma = np.zeros((n, n), dtype=int)
        Mw = np.zeros((n, n), dtype=float)
        while "EDGE_DATA_SECTION" not in line:
            line = f.readline()
        line = f.readline()
        while "-1" not in line:
            i, j = [int(x) for x in line.split()]
            Ma[i, j] = 1
            line = f.readline()
        while "EDGE_WEIGHT_SECTION" not in line:
            line = f.readline()
        for i in range(n):
            Mw[i, :] = [float(x) for x in f.readline().split()]
        while "TOUR_SECTION" not in line:
            line = f.readline()
        route = [int(x) for x in f.readline().split()]
    return Ma, Mw, route


def create_dataset(path, nmin, nmax, conn_min=1, conn_max=1, samples=1000, distances="euc_2D", seed=None):
    """dataset.py:118-143 without Concorde."""
    os.makedirs(path, exist_ok=True)
    rng = np.random.RandomState(seed) if seed is not None else np.random
    pyrng = random.Random(seed) if seed is not None else random
    for i in range(samples):
        n = pyrng.randint(nmin, nmax)
        Ma, Mw, route, _ = create_graph(n, rng.uniform(conn_min, conn_max), distances=distances, rng=rng)
        write_graph(Ma, Mw, "{}/{}.graph".format(path, i), route=route)


# ----------------------------------------------------------------------------
# instance_loader.py:29-80, vectorised; emits Incidence instead of dense EV
# ----------------------------------------------------------------------------
def create_batch(instances, dev=0.02, training_mode="relational", target_cost=None):
    n_instances = len(instances)
    n_vertices = np.array([x[0].shape[0] for x in instances], dtype=np.int64)
    srcs, dsts, ws, cs = [], [], [], []
    n_edges = np.zeros(n_instances, dtype=np.int64)
    n_acc = 0
    for i, (Ma, Mw, route) in enumerate(instances):
        n = int(n_vertices[i])
        x, y = np.nonzero(Ma)                      # row-major order, instance_loader.py:60
        n_edges[i] = len(x)
        srcs.append(x + n_acc)
        dsts.append(y + n_acc)
        ws.append(np.asarray(Mw)[x, y])
        # instance_loader.py:70: closing edge taken from route[1:]+route[1:] (kept as is)
        pairs = list(zip(route, route[1:] + route[1:]))
        cost = sum(Mw[min(a, b), max(a, b)] for (a, b) in pairs) / n
        if target_cost is None:
            cval = (1 - dev) * cost if i % 2 == 0 else (1 + dev) * cost
        else:
            cval = target_cost
        cs.append(np.full(len(x), cval, dtype=np.float64))
        n_acc += n
    cat = lambda lst, dt: (np.concatenate(lst) if lst else np.zeros(0)).astype(dt)
    EV = Incidence(cat(srcs, np.int32), cat(dsts, np.int32), n_acc)
    W = cat(ws, np.float64).reshape(-1, 1)
    C = cat(cs, np.float64).reshape(-1, 1)
    route_exists = np.array([i % 2 for i in range(n_instances)])   # instance_loader.py:50
    return EV, W, C, route_exists, n_vertices, n_edges


class InstanceLoader(object):
    """instance_loader.py:7-93."""

    def __init__(self, path):
        self.path = path
        self.filenames = [path + "/" + x for x in os.listdir(path)]
        random.shuffle(self.filenames)
        self.reset()

    def get_instances(self, n_instances):
        for _ in range(n_instances):
            Ma, Mw, route = read_graph(self.filenames[self.index])
            yield Ma, Mw, route        # two copies of every instance (instance_loader.py:21-23)
            yield Ma, Mw, route
            self.index += 1

    create_batch = staticmethod(create_batch)

    def get_batches(self, batch_size, dev):
        for _ in range(len(self.filenames) // batch_size):
            instances = list(self.get_instances(batch_size))
            yield InstanceLoader.create_batch(instances, dev=dev)

    def reset(self):
        random.shuffle(self.filenames)
        self.index = 0


# ----------------------------------------------------------------------------
# synthetic batches for tests and bench
# ----------------------------------------------------------------------------
def synth_instances(sizes, seed=42, connectivity=1.0, distances="euc_2D", two_opt_sweeps=2):
    """One instance per entry of ``sizes``; instance k is seeded ``RandomState(seed+k)``."""
    out = []
    for k, n in enumerate(sizes):
        rng = np.random.RandomState(seed + k)
        Ma, Mw, route, _ = create_graph(int(n), connectivity, distances, rng=rng, two_opt_sweeps=two_opt_sweeps)
        out.append((Ma, Mw, route))
    return out


def synth_batch(sizes, seed=42, dev=0.02, connectivity=1.0, distances="euc_2D"):
    return create_batch(synth_instances(sizes, seed, connectivity, distances), dev=dev)


def mixed_sizes(batch, nmin, nmax, seed=42):
    return [int(v) for v in np.random.RandomState(seed).randint(nmin, nmax + 1, size=batch)]
