"""The learning-rate schedule the reference names but never builds.

train_model.py:356-365 defines ``'cosine_warmup': {'type': 'custom_warmup', 'params': {'warmup_epochs': 5, 'max_lr': 1e-4,
'final_lr': 1e-7, 'total_epochs': num_epochs}}`` and steps it once per epoch (:451-452), but :376-387 has no branch that
creates a scheduler for that type, so ``-r cosine_warmup`` trains at the constant command-line rate for one epoch and
then dies with UnboundLocalError.  ``CosineWarmupLR`` is that missing object, taking exactly the config's ``params``:

    elif scheduler_config['type'] == 'custom_warmup':                      # the branch a maintainer adds after :387
        scheduler = ctk.CosineWarmupLR(optimizer, **scheduler_config['params'])

Epoch e (= number of ``scheduler.step()`` calls so far; the constructor sets epoch 0 like every torch scheduler):
    e <  warmup_epochs : lr = max_lr * (e + 1) / warmup_epochs                      (linear ramp ending at max_lr)
    e >= warmup_epochs : lr = final_lr + (max_lr - final_lr) * (1 + cos(pi * t)) / 2,
                         t = min(1, (e - warmup_epochs + 1) / (total_epochs - warmup_epochs))
so the last epoch of the run (e = total_epochs - 1) trains at final_lr and later epochs stay there.  Pure host logic on
``optimizer.param_groups``: it drives ctk.Adam and torch optimizers alike and checkpoints through ``state_dict()``.
"""
from __future__ import annotations

import math

from torch.optim.lr_scheduler import LRScheduler


class CosineWarmupLR(LRScheduler):
    def __init__(self, optimizer, warmup_epochs: int = 5, max_lr: float = 1e-4, final_lr: float = 1e-7,
                 total_epochs: int = 100, last_epoch: int = -1):
        if warmup_epochs < 0 or total_epochs <= warmup_epochs:
            raise ValueError("need 0 <= warmup_epochs < total_epochs")
        if not (max_lr > 0 and 0 <= final_lr <= max_lr):
            raise ValueError("need 0 <= final_lr <= max_lr, max_lr > 0")
        self.warmup_epochs = int(warmup_epochs)
        self.max_lr = float(max_lr)
        self.final_lr = float(final_lr)
        self.total_epochs = int(total_epochs)
        super().__init__(optimizer, last_epoch)

    def lr_at(self, epoch: int) -> float:
        """Closed form of the schedule (every param group gets the same rate, as the config has a single max_lr)."""
        if epoch < self.warmup_epochs:
            return self.max_lr * (epoch + 1) / self.warmup_epochs
        t = min(1.0, (epoch - self.warmup_epochs + 1) / (self.total_epochs - self.warmup_epochs))
        return self.final_lr + (self.max_lr - self.final_lr) * 0.5 * (1.0 + math.cos(math.pi * t))

    def get_lr(self):
        return [self.lr_at(self.last_epoch) for _ in self.optimizer.param_groups]
