#!/usr/bin/env python
"""Training driver with the reference's flags and loop (train.py:104-280) on the B200 engine.

Shows the drop-in claim end to end: the body is the reference's own control flow -- ``run_batch``
feeds ``EV, W, C, time_steps, route_exists, n_vertices, n_edges`` and fetches
``[train_step, loss, acc, predictions, TP, FP, TN, FN]`` (train.py:17-63) -- with two imports
changed.  Differences forced by this environment, not by the engine:
  * datasets come from ``instances.create_dataset`` (nearest-neighbour + 2-opt tours) because the
    exact solver the reference calls (pyconcorde, dataset.py:9-50) is not installable here;
  * checkpoints are ``model.npz`` keyed by the TF variable names instead of TF Saver files.
Extra flags: ``-samples_train/-samples_test/-batches_train/-batches_test`` bound an epoch
(reference: 2**15 / 2**10 samples, 128 / 32 batches, train.py:173-191), ``-mode`` picks the arithmetic.
"""
import argparse
import os
import random
import sys
from itertools import islice

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import tsp_gnn_b200 as tg                                   # noqa: E402   (was: import tensorflow as tf)
from tsp_gnn_b200 import build_network, InstanceLoader      # noqa: E402   (was: from model / instance_loader import)
from tsp_gnn_b200.instances import create_dataset           # noqa: E402   (was: from dataset import create_dataset)


def run_batch(sess, model, batch, batch_i, epoch_i, time_steps, train=False, verbose=True):
    """train.py:17-63."""
    EV, W, C, route_exists, n_vertices, n_edges = batch
    feed_dict = {model['EV']: EV, model['W']: W, model['C']: C, model['time_steps']: time_steps,
                 model['route_exists']: route_exists, model['n_vertices']: n_vertices, model['n_edges']: n_edges}
    if train:
        outputs = [model['train_step'], model['loss'], model['acc'], model['predictions'], model['TP'], model['FP'],
                   model['TN'], model['FN']]
    else:
        outputs = [model['loss'], model['acc'], model['predictions'], model['TP'], model['FP'], model['TN'], model['FN']]
    loss, acc, predictions, TP, FP, TN, FN = sess.run(outputs, feed_dict=feed_dict)[-7:]
    if verbose:
        print('{train_or_test} Epoch {epoch_i} Batch {batch_i}\t|\t(n,m,batch size)=({n},{m},{batch_size})\t|\t'
              '(Loss,Acc)=({loss:.4f},{acc:.4f})\t|\tAvg. (Sat,Prediction)=({avg_sat:.4f},{avg_pred:.4f})'.format(
                  train_or_test='Train' if train else 'Test', epoch_i=epoch_i, batch_i=batch_i, loss=loss, acc=acc,
                  n=np.sum(n_vertices), m=np.sum(n_edges), batch_size=n_vertices.shape[0],
                  avg_sat=np.mean(route_exists), avg_pred=np.mean(np.round(predictions))), flush=True)
    return loss, acc, np.mean(route_exists), np.mean(predictions), TP, FP, TN, FN


def summarize_epoch(epoch_i, loss, acc, sat, pred, train=False):
    """train.py:65-76."""
    print('{train_or_test} Epoch {epoch_i} Average\t|\t(Loss,Acc)=({loss:.4f},{acc:.4f})\t|\tAvg. (Sat,Pred)=({avg_sat:.4f},'
          '{avg_pred:.4f})'.format(train_or_test='Train' if train else 'Test', epoch_i=epoch_i, loss=np.mean(loss),
                                    acc=np.mean(acc), avg_sat=np.mean(sat), avg_pred=np.mean(pred)), flush=True)


def ensure_datasets(train_params, test_params, seed):
    """train.py:78-102."""
    for path, p in (('instances/train', train_params), ('instances/test', test_params)):
        if not os.path.isdir(path):
            print('{} dataset not found, creating {} instances'.format(path, p['samples']), flush=True)
            create_dataset(path, p['n_min'], p['n_max'], conn_min=p['conn_min'], conn_max=p['conn_max'],
                           samples=p['samples'], distances=p['distances'], seed=seed)


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description='TSP-GNN training on the B200 engine (flags of train.py:107-119)')
    parser.add_argument('-d', default=64, type=int, help='Embedding size for vertices and edges')
    parser.add_argument('-timesteps', default=32, type=int, help='# Timesteps')
    parser.add_argument('-dev', default=0.02, type=float, help='Target cost deviation')
    parser.add_argument('-epochs', default=10000, type=int, help='Training epochs')
    parser.add_argument('-batchsize', default=8, type=int, help='Batch size')
    parser.add_argument('-seed', type=int, default=42, help='RNG seed for Python and Numpy')
    parser.add_argument('-load_from', default=None, help='Load weights from this path')
    parser.add_argument('--save', const=True, default=False, action='store_const', help='Save model?')
    parser.add_argument('-distances', default='euc_2D', help='What type of distances? (euc_2D or random)')
    parser.add_argument('-cmin', default=1, type=float, help='Min. connectivity')
    parser.add_argument('-cmax', default=1, type=float, help='Max. connectivity')
    parser.add_argument('-samples_train', default=2 ** 15, type=int)
    parser.add_argument('-samples_test', default=2 ** 10, type=int)
    parser.add_argument('-batches_train', default=128, type=int)
    parser.add_argument('-batches_test', default=32, type=int)
    parser.add_argument('-mode', default='bf16x3', choices=['bf16x3', 'bf16', 'simt'])
    args = parser.parse_args()

    random.seed(args.seed)
    np.random.seed(args.seed)
    d, time_steps, dev, batch_size = args.d, args.timesteps, args.dev, args.batchsize
    train_params = {'n_min': 20, 'n_max': 40, 'conn_min': args.cmin, 'conn_max': args.cmax,
                    'batches_per_epoch': args.batches_train, 'samples': args.samples_train, 'distances': args.distances}
    test_params = dict(train_params, batches_per_epoch=args.batches_test, samples=args.samples_test)
    ensure_datasets(train_params, test_params, args.seed)
    train_loader = InstanceLoader('instances/train')
    test_loader = InstanceLoader('instances/test')

    print('Building model ...', flush=True)
    GNN = build_network(d, mode=args.mode)
    with tg.Session(GNN) as sess:
        print('Initializing global variables ... ', flush=True)
        sess.run(tg.global_variables_initializer(seed=args.seed))
        start_epoch = 0
        if args.load_from is not None:
            sess.load_weights(args.load_from)
            start_epoch = int(args.load_from.split('=')[-1]) if '=' in args.load_from else 0   # util.py:10
        os.makedirs('training/dev={dev}'.format(dev=dev), exist_ok=True)
        keys = ['loss', 'acc', 'sat', 'pred', 'TP', 'FP', 'TN', 'FN']
        with open('training/dev={dev}/log.dat'.format(dev=dev), 'a') as logfile:
            for epoch_i in np.arange(start_epoch, start_epoch + args.epochs):
                train_loader.reset()
                test_loader.reset()
                train_stats = {k: np.zeros(train_params['batches_per_epoch']) for k in keys}
                test_stats = {k: np.zeros(test_params['batches_per_epoch']) for k in keys}
                print('Training model...', flush=True)
                for batch_i, batch in islice(enumerate(train_loader.get_batches(batch_size, dev)),
                                             train_params['batches_per_epoch']):
                    res = run_batch(sess, GNN, batch, batch_i, epoch_i, time_steps, train=True, verbose=True)
                    for k, v in zip(keys, res):
                        train_stats[k][batch_i] = v
                summarize_epoch(epoch_i, train_stats['loss'], train_stats['acc'], train_stats['sat'], train_stats['pred'],
                                train=True)
                print('Testing model...', flush=True)
                for batch_i, batch in islice(enumerate(test_loader.get_batches(batch_size, dev)),
                                             test_params['batches_per_epoch']):
                    res = run_batch(sess, GNN, batch, batch_i, epoch_i, time_steps, train=False, verbose=True)
                    for k, v in zip(keys, res):
                        test_stats[k][batch_i] = v
                summarize_epoch(epoch_i, test_stats['loss'], test_stats['acc'], test_stats['sat'], test_stats['pred'],
                                train=False)
                savepath = 'training/dev={dev}/checkpoints/epoch={epoch}'.format(
                    dev=dev, epoch=int(round(100 * np.ceil((epoch_i + 1) / 100))))
                os.makedirs(savepath, exist_ok=True)
                if args.save:
                    sess.save_weights(savepath)
                # 17 columns of train.py:253-276 (the reference copies the train TP..FN into the test columns)
                cols = [epoch_i] + [np.mean(train_stats[k]) for k in keys] + \
                       [np.mean(test_stats[k]) for k in keys[:4]] + [np.mean(train_stats[k]) for k in keys[4:]]
                logfile.write(' '.join(str(c) for c in cols) + '\n')
                logfile.flush()
