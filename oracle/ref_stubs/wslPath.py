"""Empty stand-in so the read-only reference imports in this container (golden generation only)."""


def is_windows_path(path):
    return False


def to_posix(path):
    return path
