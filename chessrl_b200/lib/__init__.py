"""Small host-side utilities (logger)."""
