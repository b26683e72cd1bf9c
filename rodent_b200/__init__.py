"""rodent_b200 -- B200-native BVH traversal / path-tracing core behind the C ABI of
``include/rodent_b200.h`` (drop-in for the hot path of AnyDSL/rodent).

The Python layer is a thin host-side mirror of the reference's tool interfaces; all
compute happens in ``librodent_b200.so`` (hand-written sm_100a CUDA).  There is no
CPU fallback: importing :mod:`rodent_b200.lib` without the built library raises.
"""
__version__ = "0.1"
