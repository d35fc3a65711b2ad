"""Stand-in for the third-party ``pypower`` package (absent here; constants only — SURVEY.md §8c).
Data and the makeYbus formula come from rpo_b200.case14, which the KATs in tests/ pin."""
from . import idx_bus, idx_gen, ppoption, api  # noqa: F401
