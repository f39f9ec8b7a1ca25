"""State <-> index conventions and log-space helpers of flowket/exact/utils.py:18-99 (host, numpy).
Bit k of the state index <-> flattened site k (C order), bit 1 <-> spin +1."""
import numpy

fsum = numpy.sum
fdot = numpy.dot


def decimal_to_binary(decimal, num_of_bits, zero_one_base=False):
    return [(decimal >> i) & 1 if zero_one_base else 2 * ((decimal >> i) & 1) - 1 for i in range(num_of_bits)]


def binary_to_decimal(binary_digits):
    return sum(1 << i for i, d in enumerate(binary_digits) if d == 1)


def binary_array_to_decimal_array(binary_digits, out=None):
    binary_digits = numpy.asarray(binary_digits)
    n = binary_digits.shape[-1]
    weights = numpy.left_shift(numpy.int64(1), numpy.arange(n, dtype=numpy.int64))
    res = ((binary_digits == 1) * weights).sum(axis=-1)
    if out is None:
        return res.astype('int32') if n < 31 else res
    out[...] = res
    return out


def decimal_array_to_binary_array(decimal, num_of_bits, zero_one_base=False, out=None):
    decimal = numpy.asarray(decimal, dtype=numpy.int64)
    bits = (decimal[:, None] >> numpy.arange(num_of_bits, dtype=numpy.int64)) & 1
    res = bits if zero_one_base else 2 * bits - 1
    if out is None:
        return res.astype(numpy.float64)
    out[...] = res
    return out


def to_log_wave_function_vector(model, batch_size=2 ** 12, out=None):
    number_of_spins = int(numpy.prod(model.input_shape[1:]))
    num_of_states = 2 ** number_of_spins
    batch_size = min(batch_size, num_of_states)
    if out is None:
        out = numpy.zeros(shape=(num_of_states,), dtype=numpy.complex128)
    for i in range(0, num_of_states, batch_size):
        batch = decimal_array_to_binary_array(numpy.arange(i, i + batch_size), number_of_spins, False).reshape(
            (batch_size,) + tuple(model.input_shape[1:]))
        out[i:i + batch_size] = model.predict(batch, batch_size=batch_size)[:, 0]
    return out


def complex_norm_log_fsum_exp(arr):
    real_arr = numpy.real(arr)
    m = numpy.max(real_arr)
    return numpy.log(fsum(numpy.exp(real_arr - m))) + m


def log_fsum_exp(arr):
    m = numpy.max(arr)
    return numpy.log(fsum(numpy.exp(arr - m))) + m


def vector_to_machine(wave_function_vector):
    def machine(batch):
        batch = numpy.asarray(batch)
        idx = binary_array_to_decimal_array(batch.reshape(batch.shape[0], -1))
        return wave_function_vector[idx][..., numpy.newaxis]
    return machine
