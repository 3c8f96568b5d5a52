"""NEM training / validation / descriptor-export driver.

Mirrors the call surface of the reference's `ModelNetTrainer` for the VMM path
(tools/Trainer.py:14-37, 115-450): same constructor, `train_nem_mvcnn(n_epochs)`,
`update_validation_accuracy_nem(epoch)`, `save_feature_nem(epoch)`, same scalar names, log lines,
LR schedule (halved every 15 epochs) and best-accuracy checkpoints - but the per-batch loop
(Trainer.py:160-203) is ONE fused call (`NEM.run_em`), there is no `nn.DataParallel` (one process
per GPU; gradients are averaged with one all-reduce when torch.distributed is initialised), the
six per-step `add_scalar` device syncs are one host read, and per-sample `.cpu()` loops are
vectorised.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch

from .dist import GradBucket
from .features import DevicePrefetcher
from .nem_utilities import clip_gradient


class ScalarWriter:
    """tensorboardX.SummaryWriter when importable (the reference's logger, Trainer.py:9,37), else a JSON-lines
    file with the same three methods the trainer uses."""

    def __init__(self, log_dir):
        self.tb = None
        self.scalars = {}
        self.path = os.path.join(log_dir, "scalars.jsonl")
        try:
            from tensorboardX import SummaryWriter  # noqa
            self.tb = SummaryWriter(log_dir)
        except Exception:
            self.f = open(self.path, "a")

    def add_scalar(self, tag, value, step):
        value = float(value)
        if self.tb is not None:
            self.tb.add_scalar(tag, value, step)
        else:
            self.f.write(json.dumps({"tag": tag, "value": value, "step": int(step), "time": time.time()}) + "\n")
        self.scalars.setdefault(tag, []).append([time.time(), int(step), value])

    def export_scalars_to_json(self, path):
        if self.tb is not None:
            self.tb.export_scalars_to_json(path)
        else:
            with open(path, "w") as f:
                json.dump(self.scalars, f)

    def close(self):
        if self.tb is not None:
            self.tb.close()
        else:
            self.f.close()


class ModelNetTrainer(object):
    def __init__(self, model, train_loader, val_loader, optimizer, loss_fn, model_name, log_dir, num_views=12, class_num=40):
        self.class_num = class_num
        self.optimizer = optimizer
        self.train_loader = train_loader
        self.val_loader = val_loader
        self.loss_fn = loss_fn
        self.model_name = model_name
        self.log_dir = log_dir
        self.num_views = num_views
        nem, c_net = model
        self.model = (nem.cuda(), c_net.cuda())
        self.world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        self.rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        self.bucket = GradBucket(list(nem.parameters()) + list(c_net.parameters()))
        if self.world > 1:
            # the backward passes write into flat per-module arenas that are averaged in place (the head's under the EM
            # backward); any parameter outside them takes the bucket's packing path
            self.bucket.attach(nem, c_net)
        self.writer = ScalarWriter(log_dir) if (log_dir is not None and self.rank == 0) else None

    # ------------------------------------------------------------------------------------------
    def _forward(self, nem, c_net, in_data, warm_up=False):
        """The reference's per-batch loop (Trainer.py:160-200) as ONE fused call.  warm_up: `-if_decoder_warm_up 1` during
        the first 10 epochs - every iteration's residual is weighted by gamma_prior (Trainer.py:184-185)."""
        h, gamma, nem_loss, intra, inter = nem.run_em(in_data, warm_up=warm_up)
        return c_net(h, gamma), gamma, nem_loss, intra, inter

    def train_nem_mvcnn(self, n_epochs):
        nem, c_net = self.model
        best_acc, i_acc = 0, 0
        nem.train()
        c_net.train()
        for epoch in range(n_epochs):
            lr = self.optimizer.state_dict()['param_groups'][0]['lr']
            if self.writer:
                self.writer.add_scalar('params/lr', lr, epoch)
            sampler = getattr(self.train_loader, "sampler", None)
            if hasattr(sampler, "set_epoch"):
                sampler.set_epoch(epoch)                                          # DistributedSampler: a new shuffle every epoch
            warm_up = nem.if_decoder_warm_up == 1 and epoch < 10                  # Trainer.py:184
            for i, i_data in enumerate(DevicePrefetcher(self.train_loader)):     # batch i+1 is copied while step i runs
                start = time.time()
                in_data = i_data[1]
                if in_data.dim() != 3:
                    raise KeyError('wrong size for data')
                in_data = in_data.cuda(non_blocking=True)
                if in_data.dtype not in (torch.float32, torch.bfloat16):
                    in_data = in_data.float()
                target = i_data[0].cuda(non_blocking=True).long()
                self.optimizer.zero_grad()
                out_data, gamma, nem_loss, intra_loss, inter_loss = self._forward(nem, c_net, in_data, warm_up)
                closs = self.loss_fn(out_data, target)
                # Trainer.py:203 - with if_total_loss == 1 the loss is closs + the MEAN of the T per-iteration NEM losses,
                # rebuilt with torch.Tensor(...), which detaches them: only the classification loss carries gradient.
                loss = closs + (nem_loss if nem.if_total_loss == 0 else nem.iteration_losses[:, 0].mean())
                i_acc += 1
                pred = torch.max(out_data, 1)[1]
                acc = (pred == target).float().mean()
                loss.backward()
                if self.world > 1:
                    self.bucket.allreduce_mean()
                if nem.bn_init_sigma == 1:
                    clip_gradient(self.optimizer, 5)      # after the all-reduce: the single-process semantics of Trainer.py:220-221
                self.optimizer.step()
                # one device->host read for all scalars of the step
                vals = torch.stack([loss.detach(), closs.detach(), nem_loss.detach(), intra_loss.detach(),
                                    inter_loss.detach(), acc]).tolist()
                if self.writer:
                    for tag, v in zip(('train/train_loss', 'train/c_loss', 'train/nem_loss', 'train/intra_loss',
                                       'train/inter_loss', 'train/train_overall_acc'), vals):
                        self.writer.add_scalar(tag, v, i_acc)
                if self.rank == 0:
                    print('epoch %d, step %d: train_loss %.3f; train_acc %.3f. Runtime %.3f'
                          % (epoch + 1, i + 1, vals[0], vals[5], time.time() - start))
            if self.rank == 0:
                print('run evaluation...')
            start = time.time()
            with torch.no_grad():
                loss, val_overall_acc, val_mean_class_acc = self.update_validation_accuracy_nem(epoch)
            if self.writer:
                self.writer.add_scalar('val/val_mean_class_acc', val_mean_class_acc, epoch + 1)
                self.writer.add_scalar('val/val_overall_acc', val_overall_acc, epoch + 1)
                self.writer.add_scalar('val/val_loss', loss, epoch + 1)
            torch.cuda.synchronize()
            if self.rank == 0:
                print('evaluation ended... runtime: ', time.time() - start)
            if val_overall_acc > best_acc and self.rank == 0:
                best_acc = val_overall_acc
                nem.save(self.log_dir, epoch)
                c_net.save(self.log_dir, epoch)
            if epoch > 0 and (epoch + 1) % 15 == 0:
                for param_group in self.optimizer.param_groups:
                    param_group['lr'] = param_group['lr'] * 0.5
        if self.writer:
            self.writer.export_scalars_to_json(self.log_dir + "/all_scalars.json")
            self.writer.close()

    # ------------------------------------------------------------------------------------------
    def update_validation_accuracy_nem(self, epoch):
        nem, c_net = self.model
        nem.eval()
        c_net.eval()
        wrong_class = np.zeros(self.class_num)
        samples_class = np.zeros(self.class_num)
        all_loss, all_correct, all_points, nb = 0.0, 0, 0, 0
        with torch.no_grad():
            for i_data in self.val_loader:
                in_data = i_data[1].cuda(non_blocking=True)
                if in_data.dtype not in (torch.float32, torch.bfloat16):
                    in_data = in_data.float()
                target = i_data[0].cuda(non_blocking=True).long()
                out_data, gamma, _, _, _ = self._forward(nem, c_net, in_data)
                pred = torch.max(out_data, 1)[1]
                all_loss += float(self.loss_fn(out_data, target))
                nb += 1
                t = target.cpu().numpy()
                wrong = (pred != target).cpu().numpy()
                samples_class += np.bincount(t, minlength=self.class_num)
                wrong_class += np.bincount(t[wrong], minlength=self.class_num)
                all_correct += int((~wrong).sum())
                all_points += t.shape[0]
        if self.world > 1:
            # every rank saw its shard of the validation set: sum the counters before any accuracy is formed
            t = torch.tensor(np.concatenate([samples_class, wrong_class, [all_loss, nb, all_correct, all_points]]),
                             dtype=torch.float64, device='cuda')
            torch.distributed.all_reduce(t)
            t = t.cpu().numpy()
            samples_class, wrong_class = t[:self.class_num], t[self.class_num:2 * self.class_num]
            all_loss, nb, all_correct, all_points = float(t[-4]), int(t[-3]), int(t[-2]), int(t[-1])
        with np.errstate(divide='ignore', invalid='ignore'):
            each_class_acc = (samples_class - wrong_class) / samples_class
        val_mean_class_acc = float(np.nanmean(each_class_acc))
        if self.rank == 0:
            np.save(str(nem.k) + '_each_class_acc.npy', each_class_acc)
            print('Total # of test models: ', all_points)
        val_overall_acc = float(all_correct) / max(all_points, 1)
        loss = all_loss / max(nb, 1)
        if self.rank == 0:
            print('val mean class acc. : ', val_mean_class_acc)
            print('val overall acc. : ', val_overall_acc)
            print('val loss : ', loss)
        nem.train()
        c_net.train()
        return loss, val_overall_acc, val_mean_class_acc

    # ------------------------------------------------------------------------------------------
    def save_feature_nem(self, epoch, out_dir=None):
        """Descriptor export (Trainer.py:367-450): the 4096-d output of the head's second Linear per object,
        one .npy each.  Batched (the reference asserts batch size 1); files go to
        <out_dir>/<class>/<feature dir>/<name>.npy (the reference hard-codes an absolute prefix)."""
        nem, c_net = self.model
        nem.eval()
        c_net.eval()
        assert c_net.save_fea == 1, 'build the head with save_feature=1'
        out_dir = out_dir or os.path.join(self.log_dir or '.', 'descriptors')
        written = []
        with torch.no_grad():
            for i_data in self.val_loader:
                in_data = i_data[1].cuda(non_blocking=True)
                if in_data.dtype not in (torch.float32, torch.bfloat16):
                    in_data = in_data.float()
                out_data, _, _, _, _ = self._forward(nem, c_net, in_data)
                desc = out_data.cpu().numpy()
                for j, path in enumerate(i_data[2]):
                    rel = os.path.join(*path.split('/')[-3:])
                    dst = os.path.join(out_dir, rel)
                    os.makedirs(os.path.dirname(dst), exist_ok=True)
                    np.save(dst, desc[j:j + 1])          # (1, 4096) like the reference's batch-1 output
                    written.append(dst)
        return written
