"""runNNet -- training/test driver with the option names, run-directory files and checkpoint cadence of
/root/reference/ctc_fast/runNNet.py:23-237 (run(), test()).

Kept: every reference option (--layerSize --numLayers --temporalLayer --momentum --epochs --step
--anneal --reg --dataDir --alisDir --startFile --numFiles --inputDim --rawDim --outputDim --maxUttLen
--save_every --run_desc --cfg_file --test), cfg.json / params.pk (two pickles: SGD state, then the
stack) / epoch / num_files / last_cost / sentinel / train.log, resume from --cfg_file, CUDA_DEVICE.
Site constants of run_cfg.py / decoder_config.py become options: --runDir (RUN_DIR), --outDir
(likelihood output of --test).  New: --batchSize (utterances per step; 1 = reference schedule),
--maxLabels.  Launched under torchrun it trains data-parallel (one process per GPU, NCCL all-reduce).
"""
import logging
import optparse
import os
import pickle
import time
from os.path import join as pjoin

import numpy as np

import dataLoader as dl
import parallel
import nnets.brnnet as rnnet
import sgd
from run_utils import dump_config, load_config, CfgStruct, get_git_revision, get_hostname, TimeString, touch_file
from writeLikelihoods import writeLogLikes

MAX_UTT_LEN = 5000   # decoder/decoder_config.py:28 (DATASET == 'swbd'), the reference's default


def _init_distributed():
    import torch
    import torch.distributed as dist
    if 'RANK' in os.environ and 'WORLD_SIZE' in os.environ and int(os.environ['WORLD_SIZE']) > 1:
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
        if not dist.is_initialized():
            dist.init_process_group(backend='nccl')
        return dist.get_rank(), dist.get_world_size()
    if 'CUDA_DEVICE' in os.environ:                       # runNNet.py:117-120
        torch.cuda.set_device(int(os.environ['CUDA_DEVICE']))
    else:
        torch.cuda.set_device(0)
    return 0, 1


def run(args=None):
    usage = "usage : %prog [options]"
    parser = optparse.OptionParser(usage=usage)
    parser.add_option('--cfg_file', dest='cfg_file', default=None,
                      help='File with settings from previously trained net')
    parser.add_option("--test", action="store_true", dest="test", default=False)
    # Architecture
    parser.add_option("--layerSize", dest="layerSize", type="int", default=1824)
    parser.add_option("--numLayers", dest="numLayers", type="int", default=5)
    parser.add_option("--temporalLayer", dest="temporalLayer", type="int", default=3)
    # Optimization
    parser.add_option("--momentum", dest="momentum", type="float", default=0.95)
    parser.add_option("--epochs", dest="epochs", type="int", default=20)
    parser.add_option("--step", dest="step", type="float", default=1e-5)
    parser.add_option("--anneal", dest="anneal", type="float", default=1.3,
                      help="Sets (learning rate := learning rate / anneal) after each epoch.")
    parser.add_option('--reg', dest='reg', type='float', default=0.0,
                      help='lambda for L2 regularization of the weight matrices')
    parser.add_option('--batchSize', dest='batchSize', type='int', default=1,
                      help='utterances per optimisation step (1 = the reference schedule)')
    # Data
    parser.add_option("--dataDir", dest="dataDir", type="string", default='./data/')
    parser.add_option('--alisDir', dest='alisDir', type='string', default=None)
    parser.add_option('--startFile', dest='startFile', type='int', default=1, help='Start file for running testing')
    parser.add_option("--numFiles", dest="numFiles", type="int", default=384)
    parser.add_option("--inputDim", dest="inputDim", type="int", default=41 * 15)
    parser.add_option("--rawDim", dest="rawDim", type="int", default=41 * 15)
    parser.add_option("--outputDim", dest="outputDim", type="int", default=35)
    parser.add_option("--maxUttLen", dest="maxUttLen", type="int", default=MAX_UTT_LEN)
    parser.add_option("--maxLabels", dest="maxLabels", type="int", default=511)
    # Save/Load
    parser.add_option('--save_every', dest='save_every', type='int', default=10,
                      help='During training, save parameters every x number of files')
    parser.add_option('--run_desc', dest='run_desc', type='string', default='', help='Description of experiment run')
    parser.add_option('--runDir', dest='runDir', type='string', default='./runs',
                      help='parent of the run directory (RUN_DIR of the reference run_cfg.py)')
    parser.add_option('--outDir', dest='outDir', type='string', default=None,
                      help='--test: where the log-likelihood arks go')
    parser.add_option('--quiet', action='store_true', dest='quiet', default=False)

    (opts, args) = parser.parse_args(args)

    if opts.cfg_file:
        cfg = load_config(opts.cfg_file)
    else:
        cfg = vars(opts)

    rank, world = _init_distributed()

    # These config values should be updated every time
    cfg['host'] = get_hostname()
    cfg['git_rev'] = get_git_revision()
    cfg['pid'] = os.getpid()

    # Create experiment output directory
    if not opts.cfg_file:
        output_dir = pjoin(opts.runDir, str(TimeString()))
        cfg['output_dir'] = output_dir
        if rank == 0 and not os.path.exists(output_dir):
            print('Creating %s' % output_dir)
            os.makedirs(output_dir)
        opts.cfg_file = pjoin(output_dir, 'cfg.json')
    else:
        output_dir = cfg['output_dir']
    if world > 1:
        import torch.distributed as dist
        dist.barrier()

    cfg['output_dir'] = output_dir
    cfg['in_file'] = pjoin(output_dir, 'params.pk')
    cfg['out_file'] = pjoin(output_dir, 'params.pk')
    cfg['test'] = opts.test
    if opts.test:
        cfg['dataDir'] = opts.dataDir
        cfg['numFiles'] = opts.numFiles
        cfg['startFile'] = opts.startFile
        cfg['outDir'] = opts.outDir
    for key, default in (('reg', 0.0), ('batchSize', 1), ('maxLabels', 511), ('quiet', False), ('alisDir', None)):
        if key not in cfg:
            cfg[key] = default

    # Logging
    logger = logging.getLogger()
    logger.setLevel(logging.DEBUG)
    for h in [h for h in logger.handlers if getattr(h, '_ctcb', False)]:
        logger.removeHandler(h)
    handlers = [logging.FileHandler(pjoin(output_dir, 'train.log' if rank == 0 else 'train.%d.log' % rank))]
    if not cfg['quiet']:
        handlers.append(logging.StreamHandler())
    for h in handlers:
        h._ctcb = True
        logger.addHandler(h)
    logger.info('Running on %s' % cfg['host'])

    # seed for debugging, turn off when stable                       runNNet.py:113-115
    np.random.seed(33)
    import random
    random.seed(33)

    opts = CfgStruct(**cfg)

    # Testing
    if opts.test:
        test(opts)
        return

    alisDir = opts.alisDir if opts.alisDir else opts.dataDir
    loader = dl.DataLoader(opts.dataDir, opts.rawDim, opts.inputDim, alisDir)

    per_rank = parallel.per_rank_capacity(opts.batchSize, world)
    nn = rnnet.NNet(opts.inputDim, opts.outputDim, opts.layerSize, opts.numLayers, opts.maxUttLen,
                    temporalLayer=opts.temporalLayer, reg=opts.reg, maxUtts=per_rank,
                    maxLabels=min(opts.maxLabels, opts.maxUttLen))
    nn.initParams()

    SGD = sgd.SGD(nn, opts.maxUttLen, alpha=opts.step, momentum=opts.momentum, batchSize=opts.batchSize,
                  verbose=(rank == 0 and not opts.quiet))

    # Dump config
    cfg['param_count'] = nn.paramCount()
    if rank == 0:
        dump_config(cfg, opts.cfg_file)

    # Training
    epoch_file = pjoin(output_dir, 'epoch')
    if os.path.exists(epoch_file):
        start_epoch = int(open(epoch_file, 'r').read()) + 1
    else:
        start_epoch = 0

    # Load model if specified
    if os.path.exists(opts.in_file):
        with open(opts.in_file, 'rb') as fid:
            SGD.fromFile(fid)
            SGD.alpha = SGD.alpha / (opts.anneal ** start_epoch)
            nn.fromFile(fid)

    num_files_file = pjoin(output_dir, 'num_files')

    for k in range(start_epoch, opts.epochs):
        perm = np.random.permutation(opts.numFiles) + 1
        loader.loadDataFileAsynch(perm[0])

        file_start = 0
        if k == start_epoch:
            if os.path.exists(num_files_file):
                file_start = int(open(num_files_file, 'r').read().strip())
                logger.info('Starting from file %d, epoch %d' % (file_start, start_epoch))
        elif rank == 0:
            open(num_files_file, 'w').write(str(file_start))

        for i in range(file_start, perm.shape[0]):
            start = time.time()
            data_dict, alis, keys, sizes = loader.getDataAsynch()
            # Prefetch
            if i + 1 < perm.shape[0]:
                loader.loadDataFileAsynch(perm[i + 1])
            SGD.run(data_dict, alis, keys, sizes)
            end = time.time()
            logger.info('File time %f' % (end - start))

            # Save parameters and cost
            if (i + 1) % opts.save_every == 0 and rank == 0:
                logger.info('Saving parameters')
                with open(opts.out_file, 'wb') as fid:
                    SGD.toFile(fid)
                    nn.toFile(fid)
                    open(num_files_file, 'w').write('%d' % (i + 1))
                logger.info('Done saving parameters')
                if SGD.expcost:
                    with open(pjoin(output_dir, 'last_cost'), 'w') as fid:
                        if opts.reg > 0.0 and SGD.regcost:
                            fid.write(str(SGD.expcost[-1] - SGD.regcost[-1]))
                        else:
                            fid.write(str(SGD.expcost[-1]))

        if rank == 0:
            # Save epoch completed
            open(pjoin(output_dir, 'epoch'), 'w').write(str(k))
            # Save parameters for the epoch
            with open(opts.out_file + '.epoch{0:02}'.format(k), 'wb') as fid:
                SGD.toFile(fid)
                nn.toFile(fid)

        SGD.alpha = SGD.alpha / opts.anneal

    # Run now complete, touch sentinel file
    if rank == 0:
        touch_file(pjoin(output_dir, 'sentinel'))
    return SGD, nn


def test(opts):
    old_opts = CfgStruct(**load_config(opts.cfg_file))
    logger = logging.getLogger()
    logger.info('Running on %s' % get_hostname())

    with open(old_opts.in_file, 'rb') as fid:
        pickle.load(fid)  # SGD data, not needed
        alisDir = opts.alisDir if opts.alisDir else opts.dataDir
        loader = dl.DataLoader(opts.dataDir, old_opts.rawDim, old_opts.inputDim, alisDir)
        nn = rnnet.NNet(old_opts.inputDim, old_opts.outputDim, old_opts.layerSize, old_opts.numLayers,
                        old_opts.maxUttLen, temporalLayer=old_opts.temporalLayer, train=False)
        nn.initParams()
        nn.fromFile(fid)

    out_dir = opts.outDir if getattr(opts, 'outDir', None) else pjoin(opts.output_dir, 'ctc_loglikes')
    if not os.path.exists(out_dir):
        os.makedirs(out_dir)
    for i in range(opts.startFile, opts.numFiles + 1):
        writeLogLikes(loader, nn, i, out_dir, writePickle=True)


if __name__ == '__main__':
    run()
