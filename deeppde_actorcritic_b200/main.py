"""Drop-in for the reference's main.py (same flags, same JSON keys, same output files).

    python main.py --config_path=configs/lqr_d5.json [--exp_name=...] [--compute_dtype=float32|float64|config]

Writes ./logs/{exp}_config.json, {exp}_{sample}_{scheme}_{TD}_{train}.csv and ..._hist.csv exactly
as the reference does (main.py:43-68).  Under torchrun every rank computes, rank 0 writes.
"""
from __future__ import annotations

import json
import logging
import os

import numpy as np
from absl import app, flags
from absl import logging as absl_logging

from . import equation as eqn
from .config import load_config
from .solver import ActorCriticSolver

flags.DEFINE_string('config_path', 'configs/lqr_d5.json', """The path to load json file.""")
flags.DEFINE_string('exp_name', None, """The name of numerical experiments, prefix for logging""")
flags.DEFINE_string('compute_dtype', 'float32', """float32 (default), float64, or config (= net_config.dtype)""")
flags.DEFINE_string('impl', None, """tensor (tcgen05 MLP layers; default for float32) or exact (CUDA-core FMA; default for float64)""")
flags.DEFINE_integer('seed', None, """seed of weights and device sampling (the reference seeds nothing)""")
flags.DEFINE_integer('num_iterations', None, """override net_config.num_iterations""")
flags.DEFINE_string('checkpoint', None, """checkpoint file: written every --checkpoint_every iterations, resumed from if it exists""")
flags.DEFINE_integer('checkpoint_every', 1000, """iterations between checkpoints""")
FLAGS = flags.FLAGS


def _init_distributed():
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
        return dist.get_rank()
    return 0


def main(argv):
    del argv
    FLAGS.log_dir = './logs'
    if FLAGS.exp_name is None:
        FLAGS.exp_name = os.path.splitext(os.path.basename(FLAGS.config_path))[0]
    config = load_config(FLAGS.config_path)
    if FLAGS.num_iterations is not None:
        config.net_config.num_iterations = FLAGS.num_iterations
    rank = _init_distributed()
    bsde = getattr(eqn, config.eqn_config.eqn_name)(config.eqn_config)
    dim = config.eqn_config.dim
    control_dim = config.eqn_config.control_dim
    sample = config.train_config.sample_type
    scheme = config.train_config.scheme
    TD = config.train_config.TD_type
    train = config.train_config.train

    path_prefix = os.path.join(FLAGS.log_dir, FLAGS.exp_name)
    if rank == 0:
        os.makedirs(FLAGS.log_dir, exist_ok=True)
        with open('{}_config.json'.format(path_prefix), 'w') as outfile:
            json.dump(dict(config), outfile, indent=2)
    absl_logging.get_absl_handler().setFormatter(logging.Formatter('%(levelname)-6s %(message)s'))
    absl_logging.set_verbosity('info')
    logging.info('Begin to solve %s ' % config.eqn_config.eqn_name)
    solver = ActorCriticSolver(config, bsde, compute_dtype=FLAGS.compute_dtype, seed=FLAGS.seed, impl=FLAGS.impl)
    if FLAGS.checkpoint:
        solver.checkpoint_path, solver.checkpoint_every = FLAGS.checkpoint, FLAGS.checkpoint_every
        if os.path.exists(FLAGS.checkpoint):
            solver.load_checkpoint(FLAGS.checkpoint)
            logging.info('resumed from %s at iteration %d' % (FLAGS.checkpoint, solver._iter))
    training_history, x, y, true_y, z, true_z, grad_y = solver.train()
    if rank != 0:
        return
    char = sample + "_" + scheme + "_" + TD + "_" + train
    np.savetxt('{}_{}.csv'.format(path_prefix, char), training_history,
               fmt=['%d', '%.5e', '%.5e', '%.5e', '%.5e', '%.5e', '%.5e', '%.5e', '%d'], delimiter=",",
               header='step, loss_critic, loss_actor, err_value, error_value_infty, err_control, err_value_grad,error_cost2, elapsed_time',
               comments='')
    figure_data = np.concatenate([x, y, true_y, z, true_z], axis=1)
    head = ("x,") * dim + "y_NN,y_true," + ("Z_NN,") * control_dim + "z_true" + (",z_true") * (control_dim - 1)
    np.savetxt('{}_{}_hist.csv'.format(path_prefix, char), figure_data, delimiter=",", header=head, comments='')


def run():
    app.run(main)


if __name__ == '__main__':
    run()
