"""Test-infrastructure stub: stdlib stand-in for `lxml` so the reference imports offline."""
from . import etree  # noqa: F401
