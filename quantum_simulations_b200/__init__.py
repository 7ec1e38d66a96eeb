"""B200-native statevector engine behind the wenbo_engine circuit/runner surface.

Layout mirrors the reference package (wenbo_engine/): ``circuit`` (contract, fusion,
staging, pass compiler), ``kernel`` (gate library + the CUDA operator face),
``runner`` (single_node.run / collect_state), ``storage`` + ``wal`` (checkpoint role),
``csrc`` (sm_100a kernels + the C ABI in include/qsv.h, built to libqsv.so).
"""
__version__ = "0.1.0"
