"""CPU oracle for the DxMI sampler-rollout hot path.  TEST INFRASTRUCTURE ONLY.

A plain fp32 PyTorch-on-CPU restatement of the reference algorithm (swyoon/Diffusion-by-MaxEntIRL), function by
function, each citing the reference file:line it follows.  It exists to *check* the CUDA path:

  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it;
  * the product package (`diffusion_by_maxentirl_b200/`) never imports it and has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors (README.md:32-36 points at a `tests/` directory
that does not exist).  The oracle is therefore pinned against outputs of the reference itself, run in the build
container from `/root/reference` by `oracle/gen_golden.py` (committed), which (a) asserts oracle == reference to
fp32 round-off on every tensor of the path and (b) writes the small fixtures under `tests/golden/`.  The only
in-tree known-answer constant, the T=10 `user_defined_eta` comment (models/DxMI/trainer.py:147-149), is asserted
in tests/test_oracle_cpu.py.
"""
