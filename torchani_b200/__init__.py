"""torchani_b200 -- B200-native (sm_100a) energy+force hot path for ANI-style potentials.

Neighbour search -> AEV -> per-element 8-member MLP ensemble -> forces, as hand-written CUDA
behind a C-ABI (``include/ani_b200.h``), mirrored by drop-in modules with the interfaces of
``torchani.AEVComputer`` / ``torchani.nn.ANINetworks`` / ``Ensemble`` / ``torchani.neighbors``.
There is no CPU or PyTorch fallback: importing the compute modules without the built
``libani_b200.so`` raises.
"""
__version__ = "0.1.0"
