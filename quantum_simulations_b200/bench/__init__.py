"""Benchmarks of the reference's ``wenbo_engine/bench`` package, re-aimed at the GPU path.

* ``kernel``            per-gate streaming kernels and the read-only observable kernels against the HBM roofline
                        (reference: bench/kernel.py — scalar vs batched GB/s per gate)
* ``end_to_end``        ``runner.single_node.run`` / ``runner.pipeline.run`` with ``kernel="cuda"``
                        (reference: bench/end_to_end.py)
* ``matmul_vs_io``      gate kernels against the data paths that feed them: chunk files, pinned PCIe copies
                        (reference: bench/matmul_vs_io.py)
* ``io``, ``hyperparam_sweep``  chunk-store / pinned-copy bandwidth and the runner sweep (reference: bench/io.py,
                        bench/hyperparam_sweep.py)
* ``mqt_bench_runner``  MQT-Bench families: correctness + performance table (reference: bench/mqt_bench_runner.py)

The headline metric of the repository is measured by ``bench.py`` at the repo root; these are the secondary
tables.  All of them need a CUDA device (libqsv has no CPU path)."""
