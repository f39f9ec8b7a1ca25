"""A lattice operator seen through raster-flattened configurations, for 1-D machines that model a 2-D lattice site by site
in raster order (BASELINE.json configs[3]: J1J2 6x6 with ComplexValuesSimpleConvNetAutoregressive1D over 36 sites; the
composition is not in the reference).  Flattening is C order -- the same order as the device term
table's site indices -- so the device descriptor is shared and only shapes change on the host."""
import numpy as np


class FlattenedOperator(object):
    def __init__(self, operator):
        self.operator = operator
        self.lattice_shape = tuple(operator.hilbert_state_shape)
        self.hilbert_state_shape = (int(np.prod(self.lattice_shape)),)
        self.max_number_of_local_connections = operator.max_number_of_local_connections

    def device_desc(self):
        return self.operator.device_desc()

    def _to_lattice(self, sample):
        sample = np.asarray(sample) if not hasattr(sample, 'is_cuda') else sample
        return sample.reshape((sample.shape[0],) + self.lattice_shape)

    def find_conn_device(self, sample):
        conn, mel, use = self.operator.find_conn_device(self._to_lattice(sample))
        return conn.reshape(conn.shape[:2] + self.hilbert_state_shape), mel, use

    def find_conn(self, sample):
        conn, mel, use = self.operator.find_conn(self._to_lattice(sample))
        return conn.reshape(conn.shape[:2] + self.hilbert_state_shape), mel, use

    def use_state(self, state):
        return self.operator.use_state(np.asarray(state).reshape(self.lattice_shape))

    def random_states(self, num_of_states):
        return np.asarray(self.operator.random_states(num_of_states)).reshape((num_of_states,) + self.hilbert_state_shape)
