"""Import-only stub."""
