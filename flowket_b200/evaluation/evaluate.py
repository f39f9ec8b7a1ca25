"""`evaluate` / `exact_evaluate` of flowket/evaluation/evaluate.py:5-37: drive a generator for `steps` batches
with no parameter update, let the stats callbacks fill one `logs` dict per batch, return the per-key means."""
import numpy


def mean_logs(logs_arr, keys=None):
    if not logs_arr:
        return {}
    keys = list(logs_arr[0].keys()) if keys is None else list(keys)
    return {key: numpy.mean([logs[key] for logs in logs_arr]) for key in keys}


def evaluate(generator, steps, callbacks, keys_to_progress_bar_mapping=None, verbose=True):
    """Mean of every logged quantity over `steps` fresh batches (sample + local energy on the device per batch)."""
    logs_arr = []
    steps_iter = range(steps)
    progress_bar = None
    if verbose:
        try:
            import tqdm
            steps_iter = progress_bar = tqdm.trange(steps)
        except ImportError:      # the progress bar is cosmetic
            progress_bar = None
    for i in steps_iter:
        next(generator)
        logs = {}
        for callback in callbacks:
            callback.on_batch_end(i, logs)
            callback.on_epoch_end(i, logs)
        logs_arr.append(logs)
        if progress_bar is not None and keys_to_progress_bar_mapping is not None:
            shown = mean_logs(logs_arr, keys=keys_to_progress_bar_mapping)
            progress_bar.set_postfix({name: shown[key] for key, name in keys_to_progress_bar_mapping.items()})
    return mean_logs(logs_arr)


def exact_evaluate(exact_variational, callbacks):
    """One full enumeration; the callbacks see the batch index of a completed cycle."""
    exact_variational.machine_updated()
    logs = {}
    for callback in callbacks:
        callback.on_batch_end(exact_variational.num_of_batch_until_full_cycle, logs)
        callback.on_epoch_end(1, logs)
    return logs
