"""Import-only stub for the golden-vector generator."""
