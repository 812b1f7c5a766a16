"""See ``rapidfuzz/__init__.py``."""
DamerauLevenshtein = None
