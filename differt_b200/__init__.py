"""differt_b200 — B200-native DiffeRT geometric hot path."""
