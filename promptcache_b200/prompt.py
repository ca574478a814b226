"""PML prompt front-end: `<prompt schema="..."> <module arg="..."/> ... trailing text </prompt>`.

Same classes and error behaviour as reference promptcache/prompt.py (Prompt :140-210, ModuleRef :96-131,
Argument :134-137, preprocessors :54-93), written on the standard-library XML parser with a small
tolerant pre-pass (`loads_xml`) because lxml's recover mode is not available offline.
"""
from __future__ import annotations

import re
import xml.etree.ElementTree as ET
import xml.sax.saxutils
from dataclasses import dataclass
from typing import Any, Callable, List, Optional, Union

_NAME_START = re.compile(r"[A-Za-z_/!?]")
_ENTITY = re.compile(r"&(#[0-9]+|#x[0-9A-Fa-f]+|[A-Za-z][A-Za-z0-9]*);")


def loads_xml(text: str) -> ET.Element:
    """Parse XML; on failure retry after escaping stray `<` and `&` (what lxml recover=True tolerates in the
    reference's examples, e.g. raw `<` inside Python code in examples/code_generation_game.xml:167)."""
    try:
        return ET.fromstring(text)
    except ET.ParseError:
        out = []
        i, n = 0, len(text)
        while i < n:
            ch = text[i]
            if ch == "<" and not (i + 1 < n and _NAME_START.match(text[i + 1])):
                out.append("&lt;")
            elif ch == "&" and not _ENTITY.match(text, i):
                out.append("&amp;")
            else:
                out.append(ch)
            i += 1
        return ET.fromstring("".join(out))


def dumps_xml(e: ET.Element) -> bytes:
    return ET.tostring(e)


def repr_indent(obj: Any, indent: int = 1) -> str:
    return "\n".join(indent * "\t" + s for s in repr(obj).split("\n"))


def apply_preproc(text: str, preprocessors: Optional[List[Callable[[str], str]]] = None) -> str:
    for p in preprocessors or []:
        text = p(text)
    return text


def read_file(filename: str, preprocessors: Optional[List[Callable[[str], str]]] = None) -> str:
    with open(filename, "r") as f:
        return apply_preproc(f.read(), preprocessors)


def escape_xml(data: str) -> str:
    return xml.sax.saxutils.escape(data, entities={"'": "&apos;", '"': "&quot;"})


def compact_surrounding_spaces(text: str) -> str:
    """Leading and trailing whitespace runs each collapse to ONE space (prompt.py:46-47)."""
    return re.sub(r"^\s+|\s+$", " ", text)


def compact_spaces(text: str) -> str:
    return " ".join(text.split())


class Preprocessor:
    def __call__(self, prompt: str) -> str:
        raise NotImplementedError


class PreprocessorList(Preprocessor):
    def __init__(self, pre: List[Callable[[str], str]]):
        self.pre = pre

    def __call__(self, prompt: str) -> str:
        return apply_preproc(prompt, self.pre)


class CompactSpaces(Preprocessor):
    def __init__(self, only_surrounding: bool = False):
        self.only_surrounding = only_surrounding

    def __call__(self, prompt: str) -> str:
        return compact_surrounding_spaces(prompt) if self.only_surrounding else compact_spaces(prompt)


@dataclass
class Argument:
    name: str
    value: str


def _blank(s: Optional[str]) -> bool:
    return s is None or len(s.strip()) == 0


class ModuleRef:
    """A module selected by the prompt; attributes are parameter arguments, children nested selections."""

    name: str
    args: List[Argument]
    modules: List["ModuleRef"]

    def __init__(self, spec: Optional[ET.Element] = None):
        if spec is not None:
            self._from_element(spec)

    def _from_element(self, root: ET.Element) -> None:
        self.name = root.tag
        self.args = [Argument(k, v) for k, v in root.attrib.items()]
        self.modules = []
        if not _blank(root.text):
            raise ValueError("Module reference cannot have text")
        for child in root:
            self.modules.append(ModuleRef(child))
            if not _blank(child.tail):
                raise ValueError("Module reference cannot have text")

    def __repr__(self) -> str:
        args = " ".join(f"{a.name}={a.value!r}" for a in self.args)
        head = f"@{self.name}({args})" if args else f"@{self.name}"
        return "\n".join([head] + [repr_indent(m) for m in self.modules])


class Prompt(ModuleRef):
    schema: str
    text: str

    def __init__(self, spec: Union[str, ET.Element], preproc: Optional[List[Callable[[str], str]]] = None):
        super().__init__()
        self.args = []
        self.preproc = list(preproc) if preproc is not None else []
        self.text = ""
        if isinstance(spec, str):
            spec = loads_xml(apply_preproc(spec, self.preproc))
        self._from_prompt(spec)

    def _from_prompt(self, root: ET.Element) -> None:
        if root.tag != "prompt":
            raise AssertionError("root element must be <prompt>")
        self.schema = root.attrib.get("schema", "")
        self.name = self.schema
        self.modules = []
        children = list(root)
        if not children:  # text-only prompt
            if root.text is None:
                raise ValueError("Prompt cannot be empty")
            self.text = compact_surrounding_spaces(root.text)
        else:
            if not _blank(root.text):
                raise ValueError("Prompt cannot have leading text")
            for child in children:
                self.modules.append(ModuleRef(child))
                if child.tail is not None:
                    # like the reference (prompt.py:188-195) the LAST non-None tail wins
                    self.text = compact_surrounding_spaces(child.tail)
        self.text = self.text.strip()

    def add_text(self, text: str) -> None:
        self.text += apply_preproc(text, self.preproc)

    def __repr__(self) -> str:
        lines = [f"Schema: @{self.name}"] + [repr_indent(m) for m in self.modules] + ["Text: " + repr(self.text)]
        return "\n".join(lines)
