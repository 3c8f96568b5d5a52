"""Command line of the VMM / NEM trainer - the reference's `train_nem.py` flags (train_nem.py:21-64:
single-dash long options, prefix abbreviations such as `-iter` work through argparse), object
construction order and train / test modes (train_nem.py:382-496), on the CUDA library.

    python -m multi_view_sort_b200.train_nem -name NEM -train_path <dir>/ -val_path <dir>/ -cluster_n 3 -iter 10

Multi-GPU: launch with torchrun; every rank trains on its shard of each epoch (DistributedSampler) and
gradients are averaged with one all-reduce per step.
"""
from __future__ import annotations

import argparse
import itertools
import json
import os
import shutil

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from .features import KmeanImgDataset
from .nem import NEM, Multi_View_Net
from .trainer import ModelNetTrainer

# The reference's option schema (train_nem.py:21-64): names, types and defaults are the compatibility contract and are
# checked against the reference's own parser in tests/test_features.py; the descriptions are this implementation's.
# Deviation: -train_path / -val_path default to "" (the reference defaults to its authors' private mount points) and
# are validated in main().
_INT, _FLT, _STR = int, float, str
OPTIONS = [
    # (flags, type, default, description)
    (("-name", "--name"), _STR, "MVCNN_kmean", "experiment name = log / checkpoint directory"),
    (("-bs", "--batchSize"), _INT, 32, "objects per batch (per rank under torchrun)"),
    (("-num_models",), _INT, 1000, "cap on objects per class (accepted for compatibility; the feature store is read whole)"),
    (("-lr",), _FLT, 5e-5, "Adam learning rate; halved every 15 epochs"),
    (("-weight_decay",), _FLT, 0.0001, "Adam weight decay"),
    (("-cnn_name", "--cnn_name"), _STR, "vggm", "backbone that produced the view features; fixes the feature width"),
    (("-num_views",), _INT, 12, "views per object (M)"),
    (("-train_path",), _STR, "", "root of the training feature store: <root>/<class>/<feature dir>/*.npy"),
    (("-val_path",), _STR, "", "root of the validation feature store"),
    (("-fea_type",), _STR, "fc", "view feature kind; only 'fc' (vectors) is implemented"),
    (("-layer_norm",), _INT, 1, "LayerNorm inside the NEM encoder/decoder (always on in this implementation)"),
    (("-cluster_n",), _INT, 4, "latent views per object (K)"),
    (("-rnn_hidden_size",), _INT, 1024, "GRU state width (H) = width of one latent view"),
    (("-iter_num",), _INT, 15, "EM iterations per object (T); `-iter N` works as an abbreviation"),
    (("-stop_gamma_grad",), _INT, 0, "1: detach the responsibilities inside the intra-cluster loss"),
    (("-if_sort",), _INT, 1, "0 concat latent views as they are, 1 sort by score, 2 sort by responsibility spread"),
    (("-epoch",), _INT, 70, "training epochs"),
    (("-if_pool",), _INT, 0, "1: max-pool the latent views instead of concatenating them"),
    (("-train_or_test_mode",), _STR, "train", "'train', or 'test' = export descriptors with a saved model"),
    (("-modelfile",), _STR, "", "checkpoint file name inside <name>/<module>/ (default: the latest)"),
    (("-init_type",), _STR, "bin_uniform", "initial responsibilities: bin_uniform | bin_rand | rand"),
    (("-no_sig",), _INT, 1, "kept for compatibility (the reference never reads it on this path)"),
    (("-bn_init_sigma",), _INT, 0, "1: gradient clipping at 5 (the tensor-sigma branch itself is not implemented)"),
    (("-if_bn",), _INT, 0, "kept for compatibility (unused by the reference's NEM)"),
    (("-if_gamma_prior",), _INT, 1, "outer-loss variant: 0 Gaussian KL, 1 view-prior KL, 2 view-prior KL + entropy terms"),
    (("-if_decoder_warm_up",), _INT, 0, "1: weight residuals by the view prior during the first 10 epochs"),
    (("-if_total_loss",), _INT, 0, "1: log/add the mean of all T NEM losses (detached) instead of the last one"),
    (("-arch",), _INT, 0, "kept for compatibility"),
    (("-e_sigma",), _FLT, 0.1, "standard deviation of the E-step's Gaussian"),
    (("-dataset",), _STR, "MNet40", "class list: MNet40 | MNet10 | ShapeNet"),
    # ---- additions of this implementation ----
    (("-precision",), _STR, "fp32", "fp32 | bf16 | bf16_b8 | bf16_fast: tensor-core operand precision (DESIGN.md 6)"),
    (("-model_path",), _STR, "", "test mode: directory that holds <name>/NEM and <name>/MV_C_Net"),
    (("-feature_out",), _STR, "", "test mode: where the descriptors are written"),
]


def make_parser():
    ap = argparse.ArgumentParser(description="VMM / NEM trainer on the sm_100a library (the reference's train_nem.py flags)")
    for flags, typ, default, text in OPTIONS:
        ap.add_argument(*flags, type=typ, default=default, help=text)
    ap.add_argument("-no_pretraining", dest="no_pretraining", action="store_true",
                    help="kept for compatibility (no image backbone is built on this path)")
    ap.set_defaults(train=False)
    return ap


parser = make_parser()


def create_folder(log_dir):
    """train_nem.py:67-74 - an existing folder is WIPED, like the reference."""
    if not os.path.exists(log_dir):
        os.mkdir(log_dir)
    else:
        print(log_dir + '   WARNING: summary folder already exists!! It will be overwritten!!')
        shutil.rmtree(log_dir)
        os.mkdir(log_dir)


def input_dim_of(cnn_name):
    """train_nem.py:413-422."""
    if cnn_name.startswith('vgg') or cnn_name == 'alexnet':
        return 4096
    if cnn_name == 'resnet50':
        return 2048
    if cnn_name == 'googlenet':
        return 1024
    raise KeyError('the backbone is not suported yet')


def class_num_of(dataset):
    return {'MNet40': 40, 'MNet10': 10, 'ShapeNet': 55}[dataset]


def build_models(args, class_num, save_feature=0):
    """NEM then the head (the reference builds an SVCNN first only to borrow its VGG classifier layers,
    train_nem.py:408,432; the head creates equivalent layers itself)."""
    if args.fea_type != 'fc':
        raise KeyError("only 'fc' features are supported by the CUDA path")
    nem = NEM(nclasses=class_num, view_num=args.num_views, layer_norm=args.layer_norm, k=args.cluster_n,
              batch_size=args.batchSize, input_dim=input_dim_of(args.cnn_name), input_type='fc',
              rnn_hidden_size=args.rnn_hidden_size, iter_num=args.iter_num, stop_gamma_grad=args.stop_gamma_grad,
              init_type=args.init_type, no_sig=args.no_sig, bn_init_sigma=args.bn_init_sigma,
              gamma_prior_loss=args.if_gamma_prior, if_decoder_warm_up=args.if_decoder_warm_up,
              if_total_loss=args.if_total_loss, arch=args.arch, e_sigma=args.e_sigma, precision=args.precision)
    cnet_2 = Multi_View_Net(None, cnn_name=args.cnn_name, nclasses=class_num, k=args.cluster_n,
                            rnn_hidden_size=args.rnn_hidden_size, if_sort=args.if_sort, if_pooling=args.if_pool,
                            save_feature=save_feature, precision=args.precision)
    return nem, cnet_2


def main(argv=None):
    torch.random.manual_seed(10)
    torch.cuda.manual_seed_all(10)
    np.random.seed(10)
    args = parser.parse_args(argv)
    class_num = class_num_of(args.dataset)
    # reject what cannot run BEFORE create_folder() wipes an existing log directory
    if args.fea_type != 'fc':
        raise KeyError("only 'fc' features are supported by the CUDA path")
    if not args.val_path or (args.train_or_test_mode == 'train' and not args.train_path):
        raise SystemExit("-train_path / -val_path are required (the reference's defaults are its authors' private mounts)")
    input_dim_of(args.cnn_name)
    distributed = int(os.environ.get("WORLD_SIZE", "1")) > 1
    if distributed:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    rank = torch.distributed.get_rank() if distributed else 0

    def loader(ds, shuffle):
        sampler = None
        if distributed and shuffle:
            sampler = torch.utils.data.distributed.DistributedSampler(ds, shuffle=True)     # the trainer calls set_epoch()
        elif distributed:
            # evaluation: a strided shard WITHOUT the padding duplicates of DistributedSampler; the trainer sums the
            # per-rank counters (update_validation_accuracy_nem), so the metrics are those of the whole set
            world = torch.distributed.get_world_size()
            sampler = list(range(rank, len(ds), world))
        return torch.utils.data.DataLoader(ds, batch_size=args.batchSize, shuffle=(shuffle and sampler is None),
                                           sampler=sampler, num_workers=0)

    if args.train_or_test_mode == 'train':
        log_dir = args.name
        if rank == 0:
            create_folder(log_dir)
            with open(os.path.join(log_dir, 'config.json'), 'w') as f:
                json.dump(vars(args), f)
        if distributed:
            torch.distributed.barrier()
        nem, cnet_2 = build_models(args, class_num)
        nem.cuda()
        cnet_2.cuda()
        optimizer = optim.Adam(itertools.chain(nem.parameters(), cnet_2.parameters()), lr=args.lr,
                               weight_decay=args.weight_decay, betas=(0.9, 0.999))
        train_dataset = KmeanImgDataset(args.train_path, fea_type=args.fea_type, dataset=args.dataset)
        val_dataset = KmeanImgDataset(args.val_path, fea_type=args.fea_type, dataset=args.dataset)
        if rank == 0:
            print('num_train_files: ' + str(len(train_dataset.filepaths)))
            print('num_val_files: ' + str(len(val_dataset.filepaths)))
        trainer = ModelNetTrainer((nem, cnet_2), loader(train_dataset, True), loader(val_dataset, False), optimizer,
                                  nn.CrossEntropyLoss(), None, log_dir, num_views=args.num_views, class_num=class_num)
        trainer.train_nem_mvcnn(args.epoch)
    else:
        path = os.path.join(args.model_path, args.name) if args.model_path else args.name
        nem, cnet_2 = build_models(args, class_num, save_feature=1)
        nem.load(path, args.modelfile or None)
        cnet_2.load(path, args.modelfile or None)
        cnet_2.eval()
        val_dataset = KmeanImgDataset(args.val_path, fea_type=args.fea_type, dataset=args.dataset, shuffle=False)
        trainer = ModelNetTrainer((nem, cnet_2), None, loader(val_dataset, False), None, nn.CrossEntropyLoss(), None,
                                  args.name if os.path.isdir(args.name) else None, num_views=args.num_views,
                                  class_num=class_num)
        written = trainer.save_feature_nem(None, args.feature_out or None)
        if rank == 0:
            print('descriptors written: %d' % len(written))


if __name__ == '__main__':
    main()
