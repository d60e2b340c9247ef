"""b200clip: B200-native phase1_mvp query path.  Import as `b200clip` (repo root on sys.path)."""
__version__ = "0.1.0"
