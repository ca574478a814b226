"""Test-infrastructure stub for `termcolor` (generation_engine.py:118)."""


def colored(text, *a, **k):
    return text
