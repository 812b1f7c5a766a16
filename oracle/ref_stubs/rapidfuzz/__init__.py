"""Empty stand-in so the read-only reference imports in this container (golden generation only).
``process`` / ``distance.DamerauLevenshtein`` are named by ``structural_variants/sv_handler.py:5-6``; nothing the
golden vectors cover calls them."""
process = None
