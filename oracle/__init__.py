"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference algorithm for the hot path (the parity checker):
  dpm_oracle.c / index_ops.py : FPS, kNN, kNN+radius, ball query (bit-exact index ops)
  model_ref.py                : encoder + decoder forward in torch fp32

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  deeppointmap_b200/ never does, and fails loudly when
its CUDA library is missing rather than falling back to anything in here.

Pinning status: the reference ships no tests or golden vectors,
so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build
container from /root/reference (tests/test_oracle_pin.py, skipped where the
reference tree is absent) and against the fixtures those runs produced
(tests/golden/*.npz, generator: tests/golden/make_golden.py).
"""
