from easyparallellibrary_b200.communicators.collective_communicator import (
    CollectiveCommunicator, create_communicator, create_serial_communicator, create_simple_communicator,
    get_or_create, reset_registry)
from easyparallellibrary_b200.communicators.coalescing import (
    plan_buckets, estimate_split_num_for_comm, FlatBucket)
from easyparallellibrary_b200.communicators.pool import CommunicationPool
