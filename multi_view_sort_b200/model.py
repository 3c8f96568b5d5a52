"""`Model` base class - checkpoint layout compatible with the reference
(models/Model.py:7-38): ``<path>/<name>/model-<epoch:05d>.pth`` holding a plain
``state_dict``; ``load(path, modelfile=None)`` picks the lexicographically last
file when no name is given."""
import glob
import os

import torch
import torch.nn as nn


class Model(nn.Module):
    def __init__(self, name):
        super().__init__()
        self.name = name

    def save(self, path, epoch=0):
        complete_path = os.path.join(path, self.name)
        os.makedirs(complete_path, exist_ok=True)
        torch.save(self.state_dict(), os.path.join(complete_path, "model-{}.pth".format(str(epoch).zfill(5))))

    def save_results(self, path, data):
        raise NotImplementedError("Model subclass must implement this method.")

    def load(self, path, modelfile=None):
        complete_path = os.path.join(path, self.name)
        if not os.path.exists(complete_path):
            raise IOError("{} directory does not exist in {}".format(self.name, path))
        if modelfile is None:
            mf = max(glob.glob(complete_path + "/*"))
        else:
            mf = os.path.join(complete_path, modelfile)
        self.load_state_dict(torch.load(mf, map_location="cpu"))
