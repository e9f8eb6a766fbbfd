"""Stand-in for the reference package on machines without it (the GPU box): the same module layout,
class names and private attributes as wolearyc/ramannoodle for the four entry points that
``ramannoodle_b200.install()`` patches, with every method evaluated by the CPU oracle
(``oracle/numpy_port.py``).  TEST INFRASTRUCTURE ONLY — lets ``tests/test_gpu_install.py`` compare the
patched (CUDA) calls with the unpatched (oracle) ones on the same live, mutable objects."""
