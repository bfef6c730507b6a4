"""specforge_b200 — B200-native (sm_100a) EAGLE3 draft-head training step behind SpecForge's seams."""
__version__ = "0.1.0"
