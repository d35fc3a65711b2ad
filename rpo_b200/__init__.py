"""rpo_b200 — B200-native hot path of RPO (completion, GRG backward, post-correction, batched env step)."""
__version__ = "0.1.0"
