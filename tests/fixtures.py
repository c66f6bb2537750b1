"""Inputs of the reference's own loss tests, rebuilt from their description
(/root/reference/models/centernet/loss_test.py:9-49 and models/centertracker/loss_test.py:10-21): a 7x7 map with all
regression fields active, one object at (1,1), and the "perfect prediction"."""
import numpy as np


def reference_loss_fixture(track=False):
    nb_classes, H, W = 3, 7, 7
    # channel layout with every field active: 1 + 3 + 2 + 2 + 7 + 5 (+2 track) = 20 (22), y_true has the weights plane last
    obj = [0.0, 1.0, 0.0,            # class one-hot
           0.2, 0.3,                 # r_offset
           2.2, 1.1,                 # width_px, height_px
           -2.0, 1.5, 2.1, 1.2, 0.5, 1.7, 1.6,   # l_shape
           23.0, 0.12, 1.8, 1.1, 2.9]            # radial_dist, orientation, obj dims
    if track:
        obj = obj + [1.0, 2.0]       # the INTENDED order: track_offset before the weights plane
    Cp = 1 + len(obj)
    gt = np.zeros((H, W, Cp + 1), np.float64)
    gt[:, :, -1] = 1.0
    gt[1, 1, 0] = 1.0
    gt[2, 1, 0] = 0.8
    gt[1, 1, 1:Cp] = obj
    pred = gt[:, :, :Cp].copy()
    pred[2, 1, 0] = 0.0
    return nb_classes, gt[None].astype(np.float32), pred[None].astype(np.float32)
