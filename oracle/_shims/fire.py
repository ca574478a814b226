"""Test-infrastructure stub for `fire` (reference CLIs only)."""


def Fire(*a, **k):
    raise RuntimeError("fire stub")
