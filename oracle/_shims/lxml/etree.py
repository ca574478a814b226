"""Test-infrastructure stub: the subset of lxml.etree the reference calls
(promptcache/schema.py:285-286,348; promptcache/prompt.py:160-161), on stdlib ElementTree.
Well-formed XML only (no recover mode)."""
import xml.etree.ElementTree as _ET

Element = _ET.Element


class XMLParser:
    def __init__(self, recover=False, **kw):
        self.recover = recover


def fromstring(text, parser=None):
    return _ET.fromstring(text)


def tostring(e, **kw):
    return _ET.tostring(e)
