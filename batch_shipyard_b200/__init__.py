"""shipyard-b200: a Blackwell-native container-batch launcher.

Same YAML configuration surface and ``shipyard`` CLI as Azure Batch Shipyard,
re-designed for one 8xB200 NVSwitch box: the pool is the box's GPUs, a native
task runner starts one rank per GPU, and the recipes' collectives run on
hand-written sm_100a kernels (``batch_shipyard_b200.ops``).
"""
__version__ = "0.2.0"
