"""Empty stand-in: the reference's import guard (lightning_gpu.py:93-99) only asks whether a module of this name can be
found (`importlib.util.find_spec("cuquantum")`); nothing of cuQuantum is used by the B200 engine."""
