"""rl4afcs_b200 -- B200-native batched IDHP flight-control engine.

Drop-in (with a leading batch axis) for the env/agent objects of wingos80/RL4AFCS on its hot
path: ``envs.linear.env.Ce500ShortPeriod``, ``objects.RLS / Actor / Critic / IDHPsp``.  Everything
numerical runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/rl4afcs_b200.h`` (``librl4afcs_b200.so``, built in-tree by ``python -m rl4afcs_b200.build``).
There is no CPU fallback.
"""
__version__ = "0.1.0"
