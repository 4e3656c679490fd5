from easyparallellibrary_b200.profiler.flops import FlopsProfilerHook, profile_flops, measured_peaks
from easyparallellibrary_b200.profiler.memory import MemoryProfilerHook, profile_memory
from easyparallellibrary_b200.profiler.timeline import TimelineHook, kernel_table  # noqa: F401,E402
