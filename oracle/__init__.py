"""CPU oracle for the FlowKet VMC inner loop -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy + torch-CPU) of the reference's
algorithm for the hot path named in BASELINE.json `north_star`:
autoregressive sampling -> find_conn -> local energy -> gradients / SR.

Rules (enforced by tests/test_abi_and_hygiene.py::test_product_never_imports_the_oracle):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
    `--impl reference` legs may import anything from here;
  * nothing under `flowket_b200/` imports it; the product path has no CPU
    fallback and fails loudly when the CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * operators / local energy / exact-enumeration conventions: PINNED against the
    reference's own numpy code imported from /root/reference (fixtures in
    tests/golden/, generator oracle/make_golden.py) and against the
    exact-diagonalisation constants embedded in the reference's examples;
    ExactVariational, VariationalMonteCarlo + MiniBatchGenerator, the stats
    callbacks / evaluate, the MCMC diagnostics, the AutoregressiveSampler loop,
    the FastAutoregressiveSampler with its dependency graph and topologies, the
    symmetrisation ensembles, the complex SR methods and the CG solver
    are PINNED by running the reference's own classes around stand-ins for the
    TensorFlow / Keras primitives (make_golden.py exact | vmc | callbacks | mcmc |
    sampler | fast_sampler | ensembles | sr | cg).
  * network half (Keras graph; TensorFlow is not installable here): restated
    from the reference sources cited per function and PINNED against the
    reference's own machine classes run on top of oracle/tf_standin.py (an eager
    torch stand-in for the primitive TF / Keras operations; make_golden.py
    machines -> tests/golden/reference_machines.npz, 1e-10); also pinned by the reference's
    own property tests (normalisation, incremental == full) and end-to-end by
    the reference's pretrained Keras weight files
    (experiments/weights/ising_*.h5 -> published energies).  Bit-level parity
    with TF's unseeded `tf.multinomial` stream is "parity unpinned"; the
    pinned sampling contract is the explicit-uniform rule of
    deepar/samplers/autoregressive.py:37-44, under which the reference's own
    FastAutoregressiveSampler reproduces the oracle's spins exactly.
  * J1J2 connection *order* (netket, un-vendored, unpinned version): "parity
    unpinned"; values pinned by the ED constant of
    examples/j1j2_2d_monte_carlo_4.py:43.
  * real-parameter SR for ConvNetAutoregressive2D has no reference
    implementation: "parity unpinned" (algebra pinned by the complex tests).
"""
