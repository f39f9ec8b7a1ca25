"""Minimal stand-ins for the two tf.keras objects FlowKet scripts touch on this path:
`Input(shape, dtype)` and `Model(inputs, outputs)` with `.predict/.input_shape/.get_weights/...`
("machine" row).  No Keras, no TensorFlow: `predict` runs the CUDA layer program."""
import numpy as np

from . import _lib


class Input(object):
    def __init__(self, shape, dtype='int8', name=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = dtype
        self.name = name or 'input_1'


class Model(object):
    def __init__(self, inputs, outputs, name=None):
        from .machines.abstract_machine import SymbolicOutput
        if isinstance(inputs, (list, tuple)):
            inputs = inputs[0]
        if isinstance(outputs, (list, tuple)):
            outputs = outputs[0]
        if not isinstance(outputs, SymbolicOutput):
            raise TypeError('Model(outputs=...) must be a machine output such as machine.predictions')
        if outputs.machine.keras_input_layer is not inputs:
            raise ValueError('Model inputs do not match the machine input layer')
        self.input = inputs
        self.output = outputs
        self.machine = outputs.machine
        self.output_kind = outputs.kind
        self.name = name or 'model'
        self.engine = _lib.FK_ENGINE_FP32   # wave-function engine used by predict / local energy

    # ---- shape / weights (Keras names) ----------------------------------------------------------------
    @property
    def input_shape(self):
        return (None,) + self.input.shape

    @property
    def input_names(self):
        return [self.input.name]

    @property
    def output_shape(self):
        return (None, 1) if self.output_kind == 'predictions' else (None,) + self.input.shape + (2,)

    @property
    def weights(self):
        return [name for name, _, _ in self.machine.weight_specs()]

    def count_params(self):
        return self.machine.count_params()

    def get_weights(self):
        return self.machine.get_weights()

    def set_weights(self, weights):
        self.machine.set_weights(weights)

    def save_weights(self, path):
        np.savez(path, **{'w%04d' % i: w for i, w in enumerate(self.get_weights())})

    def load_weights(self, path):
        import os
        if str(path).endswith('.h5') and not os.path.exists(str(path)) and os.path.exists(str(path) + '.npz'):
            path = str(path) + '.npz'          # written by save_weights('x.h5') of this class (numpy container)
        if str(path).endswith('.h5'):
            from .utils.keras_h5 import read_keras_weights
            self.set_weights(read_keras_weights(path, self.machine.weight_specs()))
            return
        with np.load(path if str(path).endswith('.npz') else str(path) + '.npz') as f:
            if 'params' in f.files:            # a Trainer checkpoint (flat vector + optimizer slots): weights only
                flat, weights, off = f['params'], [], 0
                for _name, shp, _init in self.machine.weight_specs():
                    n = int(np.prod(shp))
                    weights.append(flat[off:off + n].reshape(shp))
                    off += n
                self.set_weights(weights)
                return
            n_w = len([k for k in f.files if k.startswith('w') and k[1:].isdigit()])
            self.set_weights([f['w%04d' % i] for i in range(n_w)])

    # ---- evaluation --------------------------------------------------------------------------------------
    def predict_device(self, x, batch_size=None):
        """torch CUDA tensor: complex64 [n,1] (predictions) or fp32 [n,*shape,2] (conditional_log_probs)."""
        return self.machine.evaluate(self.output_kind, x, batch_size=batch_size, engine=self.engine)

    def predict(self, x, batch_size=None, **_unused):
        return self.predict_device(x, batch_size=batch_size).cpu().numpy()

    __call__ = predict

    # ---- training (the three Keras calls the reference scripts make: compile, summary, fit_generator) ------------
    def compile(self, optimizer, loss=None, **_unused):
        """`loss` must be loss_for_energy_minimization (optimization/loss.py:4-5) -- the only loss of this path; its
        gradient is the hand-written device backward, not autodiff."""
        from .optimization.loss import loss_for_energy_minimization
        if loss is not None and loss is not loss_for_energy_minimization:
            raise NotImplementedError('the B200 path differentiates loss_for_energy_minimization only')
        self.optimizer = optimizer
        self.loss = loss
        self._trainer = None

    def summary(self, print_fn=print):
        specs = self.machine.weight_specs()
        print_fn('Model "%s"  input %s  output %s' % (self.name, self.input_shape, self.output_shape))
        for name, shape, _ in specs:
            print_fn('  %-44s %-20s %10d' % (name, tuple(shape), int(np.prod(shape))))
        print_fn('Total params: {:,}'.format(self.count_params()))

    def fit_generator(self, generator, steps_per_epoch=1, epochs=1, callbacks=(), initial_epoch=0, verbose=0, **kwargs):
        """Keras semantics (mini-batches per epoch); see optimizers.Trainer.fit_generator."""
        from .optimizers import Trainer
        if getattr(self, 'optimizer', None) is None:
            raise RuntimeError('compile(optimizer, loss) first')
        if self._trainer is None:
            self._trainer = Trainer(self, generator, self.optimizer)
        self._trainer.generator = generator
        return self._trainer.fit_generator(generator, steps_per_epoch=steps_per_epoch, epochs=epochs,
                                           callbacks=list(callbacks), initial_epoch=initial_epoch, verbose=verbose, **kwargs)
