"""Test-harness stub: lets the unmodified reference be imported without VTK."""
