"""Small host utilities with the reference's names (reid/utils/meters.py)."""
