"""Wall-clock checkpoints as a callback (flowket/callbacks/checkpoint.py:29-73).  The reference writes
`model.save_weights` plus a pickle of the optimizer slots; here the callback asks the Trainer it is attached
to for one .npz holding weights, optimizer slots and the sampler's Philox draw counter, so a resumed run
continues the same random stream (Trainer.save_checkpoint)."""
import time

from . import Callback


class CheckpointByTime(Callback):
    def __init__(self, filepath, save_frequency_in_minutes=30, save_weights_only=False, **kwargs):
        super(CheckpointByTime, self).__init__(**kwargs)
        self.filepath = filepath
        self.save_frequency_in_minutes = save_frequency_in_minutes
        self.save_weights_only = save_weights_only
        self.last_save_time = time.time()
        self.current_epoch = 0
        self.trainer = None     # set by Trainer.fit
        self.saves = 0

    def set_trainer(self, trainer):
        self.trainer = trainer

    def _save(self, logs):
        path = self.filepath.format(**dict(logs or {}, epoch=self.current_epoch))
        if self.save_weights_only or self.trainer is None:
            self.model.save_weights(path)
        else:
            self.trainer.save_checkpoint(path)
        self.saves += 1
        self.last_save_time = time.time()

    def on_epoch_begin(self, epoch, logs=None):
        self.current_epoch = epoch

    def on_batch_end(self, batch, logs=None):
        if time.time() - self.last_save_time >= self.save_frequency_in_minutes * 60:
            self._save(logs)

    def on_train_end(self, logs=None):
        self.current_epoch += 1
        self._save(logs)
