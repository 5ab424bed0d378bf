"""
Learning-rate schedules of the reference (/root/reference/lr.py:11-42): a scalar function of the training progress
`x` in epochs.  `GSSupervised.set_progress` (models.py:93-95) evaluates the schedule the constructor named and writes the
value into every `param_group` of the model's optimiser through `LRSchedule.set_lr` -- host-side scalar maths, mirrored
here with the same names, arguments and defaults so `lr_schedule='constant' | 'linear' | 'cyclical' | 'step'` select the
same curves.  (`step` takes no `lr_init`, so naming it in the constructor fails in the reference too: models.py:66 always
binds `lr_init`.)
"""

import math


class LRSchedule(object):

    @staticmethod
    def set_lr(optimizer, lr):
        for group in optimizer.param_groups:
            group['lr'] = lr

    @staticmethod
    def constant(x, lr_init=0.1, epochs=1):
        return lr_init

    @staticmethod
    def step(x, breaks=(150, 250)):
        rates = (0.1, 0.01, 0.001)
        return rates[sum(1 for edge in breaks[:2] if x >= edge)]

    @staticmethod
    def linear(x, lr_init=0.1, epochs=1):
        return lr_init * float(epochs - x) / epochs

    @staticmethod
    def cyclical(x, lr_init=0.1, epochs=1):
        # warm-up epoch at a fixed small rate, then a saw-tooth inside every epoch under a linear decay over the epochs
        if x < 1:
            return 0.05
        return lr_init * (1 - x % 1) * (epochs - math.floor(x)) / epochs
