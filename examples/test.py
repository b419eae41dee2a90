#!/usr/bin/env python
"""Evaluation driver with the reference's flags (test.py:11-60) on the B200 engine: loads a checkpoint,
runs every instance as a batch of its two +-dev copies, prints the mean statistics."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import tsp_gnn_b200 as tg                                   # noqa: E402
from tsp_gnn_b200 import build_network, InstanceLoader      # noqa: E402
from train import run_batch, summarize_epoch                # noqa: E402   (test.py:9 does the same)

if __name__ == '__main__':
    parser = argparse.ArgumentParser(description='TSP-GNN evaluation on the B200 engine (flags of test.py:14-19)')
    parser.add_argument('-d', default=64, type=int, help='Embedding size for vertices and edges')
    parser.add_argument('-time_steps', default=32, type=int, help='# Timesteps')
    parser.add_argument('-dev', default=0.02, type=float, help='Target cost deviation')
    parser.add_argument('-instances', default='instances/test', help='Path for the test instances')
    parser.add_argument('-checkpoint', default='training/dev=0.02/checkpoints/epoch=100',
                        help='Path for the checkpoint of the trained model')
    args = parser.parse_args()
    loader = InstanceLoader(args.instances)
    print('Building model ...', flush=True)
    GNN = build_network(args.d)
    with tg.Session(GNN) as sess:
        sess.load_weights(args.checkpoint)                   # raises 'Path does not exist!' like util.py:20
        n_instances = len(loader.filenames)
        keys = ['loss', 'acc', 'sat', 'pred', 'TP', 'FP', 'TN', 'FN']
        stats = {k: np.zeros(n_instances) for k in keys}
        for batch_i, batch in enumerate(loader.get_batches(1, args.dev)):
            res = run_batch(sess, GNN, batch, batch_i, 0, args.time_steps, train=False, verbose=True)
            for k, v in zip(keys, res):
                stats[k][batch_i] = v
        summarize_epoch(0, stats['loss'], stats['acc'], stats['sat'], stats['pred'], train=False)
