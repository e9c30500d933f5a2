"""Test-harness stub (viewer is never used on the hot path)."""
Viewer = None
