"""laser-polio per-tick agent update, B200-native."""
__version__ = "0.1.0"
