"""Empty stand-in so the read-only reference imports in this container (golden generation only)."""
