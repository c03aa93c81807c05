"""Drop-in mirrors of the `skoots.lib` callables on the instance-assembly path.

Same names, argument meaning, return conventions and error behaviour as the reference
(SURVEY.md §8b); the arithmetic runs in libskoots_b200.so on the tensor's CUDA device.
"""
