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

parser = argparse.ArgumentParser()
parser.add_argument("-name", "--name", type=str, help="Name of the experiment", default="MVCNN_kmean")
parser.add_argument("-bs", "--batchSize", type=int, help="Batch size for the second stage", default=32)
parser.add_argument("-num_models", type=int, help="number of models per class", default=1000)
parser.add_argument("-lr", type=float, help="learning rate", default=5e-5)
parser.add_argument("-weight_decay", type=float, help="weight decay", default=0.0001)
parser.add_argument("-no_pretraining", dest='no_pretraining', action='store_true')
parser.add_argument("-cnn_name", "--cnn_name", type=str, help="cnn model name", default="vggm")
parser.add_argument("-num_views", type=int, help="number of views", default=12)
parser.set_defaults(train=False)
parser.add_argument("-train_path", type=str, default="/mnt/cloud_disk/yw/MVCNN_features_vggm_pretrained_imagenet_shapenetcore_train/")
parser.add_argument("-val_path", type=str, default="/mnt/cloud_disk/yw/MVCNN_features_vggm_pretrained_imagenet_shapenetcore_test/")
parser.add_argument("-fea_type", type=str, help="feature type for NEM, fc or conv", default='fc')
parser.add_argument("-layer_norm", type=int, help="use layernorm for NEM or not", default=1)
parser.add_argument("-cluster_n", type=int, help="cluster number", default=4)
parser.add_argument("-rnn_hidden_size", type=int, help="rnn hidden size", default=1024)
parser.add_argument("-iter_num", type=int, help="iteration number for EM", default=15)
parser.add_argument("-stop_gamma_grad", type=int, help="stop the gradient propagation from gamma", default=0)
parser.add_argument("-if_sort", type=int, help="if sort the disentangled views or not ", default=1)
parser.add_argument("-epoch", type=int, help="epoch number", default=70)
parser.add_argument("-if_pool", type=int, help="use pooling rather than concatenation for multi-view", default=0)
parser.add_argument("-train_or_test_mode", type=str, help="train or test", default='train')
parser.add_argument("-modelfile", type=str, help="model file name", default='')
parser.add_argument("-init_type", type=str, help="initializetion strategy for NEM", default='bin_uniform')
parser.add_argument("-no_sig", type=int, help="no sigmoid after rnn-nem", default=1)
parser.add_argument("-bn_init_sigma", type=int, help="use minibatch data to init sigma", default=0)
parser.add_argument("-if_bn", type=int, help="use bn layers", default=0)
parser.add_argument("-if_gamma_prior", type=int, help="if use gamma prior loss", default=1)
parser.add_argument("-if_decoder_warm_up", type=int, help="if warm up decoder", default=0)
parser.add_argument("-if_total_loss", type=int, help="if use loss from each time stamp", default=0)
parser.add_argument("-arch", type=int, help="net architecture", default=0)
parser.add_argument("-e_sigma", type=float, help="net architecture", default=0.1)
parser.add_argument("-dataset", type=str, help="dataset to use, MNet40, MNet10, ShapeNet are supported", default='MNet40')
# additions of this implementation
parser.add_argument("-precision", type=str, default="fp32", help="fp32 | bf16 | bf16_h | bf16_fast (tensor-core operand precision, DESIGN.md 6)")
parser.add_argument("-model_path", type=str, default="", help="test mode: directory holding <name>/NEM and <name>/MV_C_Net "
                    "(the reference hard-codes /mnt/cloud_disk/huangjj/exp_mvcnn/)")
parser.add_argument("-feature_out", type=str, default="", help="test mode: where the descriptors are written")


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
    distributed = int(os.environ.get("WORLD_SIZE", "1")) > 1
    if distributed:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    rank = torch.distributed.get_rank() if distributed else 0

    def loader(ds, shuffle):
        sampler = torch.utils.data.distributed.DistributedSampler(ds, shuffle=shuffle) if distributed else None
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
