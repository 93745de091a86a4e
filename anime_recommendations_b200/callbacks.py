"""The three Keras callbacks neural_network.py:184-208 passes to fit(), same names and kwargs."""
from __future__ import annotations

import numpy as np


class Callback:
    def set_model(self, model):
        self.model = model

    def on_train_begin(self):
        pass

    def on_train_end(self):
        pass

    def on_epoch_begin(self, epoch):
        pass

    def on_epoch_end(self, epoch, logs):
        pass


class LearningRateScheduler(Callback):
    """tfkc.LearningRateScheduler(lambda epoch: lrfn(epoch)) -- neural_network.py:184-186."""

    def __init__(self, schedule, verbose=0):
        self.schedule, self.verbose = schedule, verbose

    def on_epoch_begin(self, epoch):
        lr = float(self.schedule(epoch))
        self.model.lr = lr
        if self.verbose:
            print("Epoch %d: LearningRateScheduler setting learning rate to %s." % (epoch + 1, lr))


def _improved(mode, cur, best):
    return cur < best if mode == "min" else cur > best


class ModelCheckpoint(Callback):
    """tfkc.ModelCheckpoint(filepath, save_weights_only, monitor, mode, save_best_only, save_freq='epoch')
    -- neural_network.py:188-196."""

    def __init__(self, filepath, save_weights_only=True, monitor="val_loss", save_freq="epoch", mode="min",
                 save_best_only=True, verbose=0, options=None):
        if save_freq != "epoch":
            raise ValueError("only save_freq='epoch' (config.yaml:72) is supported")
        if mode == "auto":
            mode = "max" if "acc" in monitor else "min"
        self.filepath, self.save_weights_only, self.monitor = filepath, save_weights_only, monitor
        self.mode, self.save_best_only, self.verbose = mode, save_best_only, verbose
        self.best = np.inf if mode == "min" else -np.inf

    def on_epoch_end(self, epoch, logs):
        cur = logs.get(self.monitor)
        if self.save_best_only:
            if cur is None or not _improved(self.mode, cur, self.best):
                return
            self.best = cur
        if self.save_weights_only:
            self.model.save_weights(self.filepath)
        else:
            self.model.save(self.filepath)
        if self.verbose:
            print("Epoch %d: %s improved to %.5f, saving model to %s" % (epoch + 1, self.monitor, cur, self.filepath))


class EarlyStopping(Callback):
    """tfkc.EarlyStopping(patience=3, monitor, mode, restore_best_weights=True) -- neural_network.py:198-201.
    Keras 2.12 semantics: weights are restored only when the callback itself stops the run."""

    def __init__(self, monitor="val_loss", min_delta=0, patience=0, verbose=0, mode="min", baseline=None,
                 restore_best_weights=False):
        if mode == "auto":
            mode = "max" if "acc" in monitor else "min"
        self.monitor, self.patience, self.mode = monitor, patience, mode
        self.min_delta = abs(min_delta)
        self.restore_best_weights, self.verbose = restore_best_weights, verbose

    def on_train_begin(self):
        self.wait, self.stopped_epoch, self.best_epoch = 0, 0, 0
        self.best = np.inf if self.mode == "min" else -np.inf
        self.best_weights = None

    def on_epoch_end(self, epoch, logs):
        cur = logs.get(self.monitor)
        if cur is None:
            return
        if self.restore_best_weights and self.best_weights is None:
            self.best_weights = self.model._device_weights()
        self.wait += 1
        delta = self.min_delta if self.mode == "min" else -self.min_delta
        if _improved(self.mode, cur + delta, self.best):
            self.best, self.best_epoch, self.wait = cur, epoch, 0
            if self.restore_best_weights:
                self.best_weights = self.model._device_weights()
            return
        if self.wait >= self.patience and epoch > 0:
            self.stopped_epoch = epoch
            self.model.stop_training = True
            if self.restore_best_weights and self.best_weights is not None:
                self.model._restore_device_weights(self.best_weights)
