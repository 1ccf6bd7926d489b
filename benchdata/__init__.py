"""Synthetic benchmark inputs (no query logic)."""
