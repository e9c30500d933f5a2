"""Test-harness stub: CPD is out of scope and never called by the oracle."""
