"""Framework-wide constants.

Parity: the reference keeps its constants in ``epl/utils/constant.py``; only
the ones that still mean something for an eager, SPMD, one-process-per-GPU
runtime are kept here (no TF op-type deny lists, no name-prefix formats for
cloned sub-graphs).
"""

# reduction of gradients over replicas / micro-batches (reference constant.py:75-76)
REDUCE_MEAN = "mean"
REDUCE_SUM = "sum"
REDUCE_METHODS = (REDUCE_MEAN, REDUCE_SUM)

# pipeline schedule policies (reference scheduler.py:120-131)
SCHEDULE_PREFER_FORWARD = "preferforward"
SCHEDULE_PREFER_BACKWARD = "preferbackward"
SCHEDULE_PREFER_BACKWARD_OPT = "preferbackwardoptimizer"
DEFAULT_PIPELINE_STRATEGY = "PreferBackward"

# gradient checkpoint (reference constant.py:92-97)
GC_COLLECTION = "collection"
GC_AUTO = "auto"
GC_COLLECTION_NAME = "checkpoints"

# ZeRO levels.  v0/v1 are what the reference ships (config.py:132-137); v2/v3
# are B200 extensions (flat-buffer sharding over NVSwitch makes them cheap).
ZERO_LEVELS = ("", "v0", "v1", "v2", "v3")
OFFLOAD_LEVELS = ("", "v0")
AMP_LEVELS = ("", "o1", "bf16", "fp8")

# broadcast / coalescing defaults (reference constant.py:81-82)
SERIAL_COMM_MAX_SPLITS = 60
COMM_SPLIT_BYTES = 32 << 20

# auto-stage policies (reference constant.py:126-128)
STAGE_POLICY_BALANCE_OP_NUM = "balance_op_num"
STAGE_POLICY_REPEATED_LAYERS = "repeated_layers"
STAGE_POLICY_HEURISTIC = "heuristic"
MIN_REPEAT_BLOCKS = 4

# MoE: dispatch happens before the first expert contraction, combine before the
# third (reference constant.py:105-106).
MOE_EINSUMS_PER_LAYER = 3

ENV_TF_CONFIG = "TF_CONFIG"
ENV_PREFIX = "EPL_"

# B200 hardware model used by planners and roofline reports.
B200_NUM_SMS = 148
B200_HBM_BYTES = 180 * (1 << 30)
B200_L2_BYTES = 126 * (1 << 20)
NVLINK_GBS_PER_DIR = 900.0
NVLINK_GBS_MEASURED = 770.0
