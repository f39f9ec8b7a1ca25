"""CPU tier: the C-ABI library loads and exports every symbol include/flowket_b200.h declares (no compute
calls), the product never imports the oracle, and there is no CPU fallback."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'flowket_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(fk_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from flowket_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    symbols = _header_symbols()
    assert len(symbols) >= 20
    for name in symbols:
        assert hasattr(lib, name), 'libflowket_b200.so does not export %s' % name
    assert set(symbols) == set(_lib.SIGNATURES), 'ctypes signature table out of sync with the header'
    assert _lib.load().fk_version() >= 100


def test_library_is_sm100a_with_tcgen05():
    """the shipped binary contains sm_100a SASS with tcgen05 MMAs, TMEM loads and bulk async copies"""
    from flowket_b200 import _lib
    out = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    for mnemonic in ('UTCHMMA', 'LDTM', 'UBLKCP'):
        assert mnemonic in out, mnemonic


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'flowket_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), os.path.join(dirpath, f)
                if f.endswith('.py'):   # (C++ comments cite reference file:line; Python must never read the tree)
                    assert '/root/reference' not in text, os.path.join(dirpath, f)


def test_no_cpu_fallback():
    """without a CUDA device the product path fails loudly instead of computing on the host"""
    import torch
    if torch.cuda.is_available():
        pytest.skip('needs a CPU-only box')
    from flowket_b200 import Input, Model, FlowketB200Error
    from flowket_b200.machines import ConvNetAutoregressive2D
    from flowket_b200.operators import Heisenberg
    from flowket_b200.samplers import FastAutoregressiveSampler
    inp = Input(shape=(4, 4))
    machine = ConvNetAutoregressive2D(inp, depth=3, num_of_channels=8)
    model = Model(inp, machine.predictions)
    with pytest.raises(FlowketB200Error):
        model.predict([[[1] * 4] * 4])
    with pytest.raises(FlowketB200Error):
        next(FastAutoregressiveSampler(Model(inp, machine.conditional_log_probs), 4))
    with pytest.raises(FlowketB200Error):
        Heisenberg(hilbert_state_shape=[4, 4], pbc=False).find_conn([[[1] * 4] * 4])


def test_missing_library_is_reported(tmp_path):
    from flowket_b200 import _lib
    with pytest.raises(_lib.FlowketB200Error):
        _lib.load(str(tmp_path / 'does_not_exist.so'))


def test_integration_guide_calls_match_the_abi():
    """every `lib.fk_*(...)` call shown in INTEGRATION.md names a function of include/flowket_b200.h and passes as many
    arguments as its ctypes signature (the binding a maintainer would copy must not rot)"""
    import re
    from flowket_b200 import _lib
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    header = open(os.path.join(ROOT, 'include', 'flowket_b200.h')).read()
    seen = 0
    for m in re.finditer(r'lib\.(fk_\w+)\(', text):
        name, i, depth, args, cur = m.group(1), m.end(), 1, [], ''
        while i < len(text):
            c = text[i]
            if c in '([{':
                depth += 1
            if c in ')]}':
                depth -= 1
                if depth == 0:
                    break
            if c == ',' and depth == 1:
                args.append(cur)
                cur = ''
            else:
                cur += c
            i += 1
        if cur.strip():
            args.append(cur)
        assert name in _lib.SIGNATURES and re.search(r'\b%s\s*\(' % name, header), name
        assert len(args) == len(_lib.SIGNATURES[name][1]), (name, len(args), len(_lib.SIGNATURES[name][1]))
        seen += 1
    assert seen >= 12


def test_hot_kernels_keep_off_the_local_memory_stack():
    """At a ~223 KB shared-memory carve-out the L1 that backs the local-memory stack is nearly gone: by-value descriptor copies
    with run-time indexing cost the Jacobian rows kernel 36 ms per step.  The kernels that were cleaned
    must stay clean, and every tensor-core kernel must keep issuing tcgen05 / bulk-copy instructions (SASS of the built library)."""
    import importlib.util
    import shutil
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump not on PATH')
    spec = importlib.util.spec_from_file_location('sass_histogram', os.path.join(ROOT, 'profiles', 'sass_histogram.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    counts, total, names = mod.collect()
    by_name = {}
    for k, n in names.items():
        by_name.setdefault(n.split('(')[0].replace('void ', ''), []).append(k)

    def one(name):
        ks = by_name.get(name)
        assert ks, (name, sorted(by_name)[:50])
        return counts[ks[0]]

    for name, limit in (('fk::tc_dw_rows_kernel', 8), ('fk::tc_dw_kernel', 0), ('fk::tc_sample_kernel', 0), ('fk::gram2_kernel', 2)):
        c = one(name)
        assert c['LDL'] + c['STL'] <= limit, (name, c['LDL'], c['STL'])
    for name in ('fk::tc_dw_rows_kernel', 'fk::tc_dw_kernel', 'fk::tc_sample_kernel', 'fk::gram2_kernel', 'fk::tcx_forward_kernel',
                 'fk::tc_backward_kernel', 'fk::gram_tc_kernel'):
        c = one(name)
        assert c['UTCHMMA'] > 0 and c['HMMA'] == 0, (name, dict(c))
    assert one('fk::gram2_kernel')['UTMALDG'] > 0          # TMA tensor-map loads
    assert one('fk::tcx_forward_kernel')['UBLKCP'] > 0      # bulk copies of the weight images
