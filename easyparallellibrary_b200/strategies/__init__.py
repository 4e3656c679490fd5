from easyparallellibrary_b200.strategies.base import ParallelStrategy, Replicate, Split, replicate, split
from easyparallellibrary_b200.strategies.context import StrategyContext
