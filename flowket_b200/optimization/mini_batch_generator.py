"""Mini-batch bookkeeping between a batch producer (`next_batch() -> (configurations, loss coefficients)`) and the
optimiser loop -- the protocol of flowket/optimization/mini_batch_generator.py:5-45, host only.

A batch of `batch_size` samples is handed out in consecutive windows of `mini_batch_size`; a new batch is drawn whenever the
next window would run past the end (so a ragged tail, batch % mini_batch, is never served -- the reference's behaviour);
`update_params_frequency` = windows per batch is what the accumulate-gradient optimiser steps on."""
import math


class MiniBatchGenerator(object):
    def __init__(self, batch_size, mini_batch_size):
        self._window = None            # (x, y) of the batch being served
        self._cursor = 0
        self.set_batch_size(batch_size, mini_batch_size)

    def next_batch(self):
        raise NotImplementedError('subclasses produce (batch, loss coefficients) here')

    def set_batch_size(self, batch_size, mini_batch_size=None):
        self.batch_size = batch_size
        self.mini_batch_size = batch_size if mini_batch_size is None else min(mini_batch_size, batch_size)
        self._window = None            # forces a fresh batch on the next request
        self.update_params_frequency = int(math.ceil(self.batch_size / float(self.mini_batch_size)))
        return self.update_params_frequency

    def _exhausted(self):
        return self._window is None or self._cursor + self.mini_batch_size > self.batch_size

    def next_mini_batch_size(self):
        if self._exhausted():
            self._window = self.next_batch()
            self._cursor = 0
        x, y = self._window
        lo, self._cursor = self._cursor, self._cursor + self.mini_batch_size
        return x[lo:self._cursor, ...], y[lo:self._cursor]

    def __next__(self):
        return self.next_mini_batch_size()

    next = __next__

    def __iter__(self):
        return self

    def to_generator(self):
        """a plain Python generator over the same stream (what the reference scripts pass to fit_generator)"""
        while True:
            yield self.next_mini_batch_size()
