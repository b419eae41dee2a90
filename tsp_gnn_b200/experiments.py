"""Inference loops of the reference's experiment scripts on top of ``Session``.

``get_cost`` mirrors experiments/binary_search.py:13-77: a binary search on the target cost C that
repeatedly evaluates ``predictions`` on the SAME graph.  ``Session.run`` recognises the unchanged
incidence and skips ``tspgnn_plan``; every probe then costs E_init + the timestep loop + read-out.
"""
from itertools import islice

import numpy as np

from .instances import create_batch


def cost_bounds(Mw):
    """binary_search.py:23-34: sum of the n lightest / heaviest entries of triu and tril (zeros of the
    other triangle included, as in the reference), normalised by n."""
    n = Mw.shape[0]
    wmin = np.minimum(np.sum(np.sort(np.triu(Mw).flatten())[:n]), np.sum(np.sort(np.tril(Mw).flatten())[:n]))
    wmax = np.maximum(np.sum(np.sort(np.triu(Mw).flatten())[-n:]), np.sum(np.sort(np.tril(Mw).flatten())[-n:]))
    return wmin / n, wmax / n


def get_cost(sess, model, instance, time_steps, threshold=0.5, stopping_delta=0.01, max_iterations=64):
    """Returns (wpred, pred, route_cost, iterations) like binary_search.py:13-77."""
    Ma, Mw, route = instance
    n = Ma.shape[0]
    wmin, wmax = cost_bounds(Mw)
    wpred = (wmin + wmax) / 2
    route = list(route)
    # binary_search.py:40 closes the tour correctly (route[1:] + route[:1]), unlike instance_loader.py:70
    route_cost = sum(Mw[min(i, j), max(i, j)] for i, j in zip(route, route[1:] + route[:1])) / n
    EV, W, _, route_exists, n_vertices, n_edges = create_batch([(Ma, Mw, route)], target_cost=wpred)
    m = int(np.sum(n_edges))
    C = np.ones((m, 1), dtype=np.float32)
    feed = {model["EV"]: EV, model["W"]: W, model["C"]: None, model["time_steps"]: time_steps,
            model["route_exists"]: route_exists, model["n_vertices"]: n_vertices, model["n_edges"]: n_edges}
    iterations, pred = 0, None
    while (wmin < wpred * (1 - stopping_delta) or wpred * (1 + stopping_delta) < wmax) and iterations < max_iterations:
        feed[model["C"]] = C * wpred
        pred = float(np.asarray(sess.run(model["predictions"], feed_dict=feed)).reshape(-1)[0])
        if pred < threshold:
            wmin = wpred
        else:
            wmax = wpred
        wpred = (wmax + wmin) / 2
        iterations += 1
    return wpred, pred, route_cost, iterations


def get_accuracy(sess, model, batch, time_steps):
    """experiments/test_varying_sizes.py:20-38 / test_varying_dev.py:20-38: ``acc`` of one batch."""
    EV, W, C, route_exists, n_vertices, n_edges = batch
    feed = {model["EV"]: EV, model["W"]: W, model["C"]: C, model["time_steps"]: time_steps,
            model["route_exists"]: route_exists, model["n_vertices"]: n_vertices, model["n_edges"]: n_edges}
    return float(np.mean(sess.run(model["acc"], feed_dict=feed)))


def get_accuracy_tp(sess, model, batch, time_steps):
    """experiments/test_varying_dev.py:21-38: that script fetches ``TP`` and divides by the batch size
    (so it reports the share of instances that are positive AND predicted right, at most 0.5 for the
    reference's half-positive batches) instead of ``acc``."""
    EV, W, C, route_exists, n_vertices, n_edges = batch
    feed = {model["EV"]: EV, model["W"]: W, model["C"]: C, model["time_steps"]: time_steps,
            model["route_exists"]: route_exists, model["n_vertices"]: n_vertices, model["n_edges"]: n_edges}
    return float(np.mean(sess.run(model["TP"], feed_dict=feed) / len(n_vertices)))


def accuracy_sweep(sess, model, loaders, devs, time_steps, batch_size=16, n_batches=64, metric="acc"):
    """The sweeps behind figures/test_varying_sizes.png and test_varying_dev.png
    (test_varying_sizes.py:82-113, test_varying_dev.py:81-96): mean accuracy over ``n_batches`` batches
    for every (key, deviation).  ``loaders`` maps a key (e.g. the instance size n) to an InstanceLoader.
    Returns {(key, dev): accuracy}; plotting and the result files stay with the caller."""
    out = {}
    for key, loader in loaders.items():
        for dev in devs:
            loader.reset()
            fn = get_accuracy_tp if metric == "TP" else get_accuracy     # "TP": test_varying_dev.py's variant
            accs = [fn(sess, model, batch, time_steps)
                    for batch in islice(loader.get_batches(batch_size, dev), n_batches)]
            out[(key, dev)] = float(np.mean(accs)) if accs else float("nan")
    return out
