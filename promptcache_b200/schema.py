"""PML schema front-end: XML -> Schema/Module/UnionModule/Parameter/TokenSequence with position offsets.

This is the PRODUCER of the offsets the splice consumes.  It keeps the class names, public methods and the
position-assignment rules of reference promptcache/schema.py:
  * every text run becomes a TokenSequence at the running offset (:326-331, :369-374); `max_tokens` keeps the
    head and tail halves of long runs (:167-168);
  * `<parameter length=n>` reserves n positions, scaffold tokens first, unk padding after (:115-141);
  * every member of a `<union>` starts at the union's offset and the union is as long as its longest member
    (:200-225) -- this is what makes cached position ids discontiguous;
  * a Scaffold is a union-free selection: one member per union on the requested paths (:449-481).
"""
from __future__ import annotations

import pathlib
import re
import xml.etree.ElementTree as ET
from typing import Any, List, Optional, Union

from .prompt import compact_surrounding_spaces, dumps_xml, loads_xml, repr_indent

_VALID_NAME = re.compile(r"^[a-zA-Z_][a-zA-Z0-9_\-.]*$")


def is_valid_xml_element_name(name: str) -> bool:
    return bool(_VALID_NAME.fullmatch(name))


def trim_with_padding(text: str, padding: int = 1) -> str:
    pad = " " * padding
    return pad + text.strip() + pad


class Path:
    """Slash-separated module path (schema.py:32-73)."""

    def __init__(self, path: Optional[Union[str, List[str]]] = None):
        if path is None:
            path = []
        if isinstance(path, str):
            if "/" in path:
                path = [s.strip() for s in path.split("/")]
            else:
                path = [path] if path else []
        self.path: List[str] = list(path)

    def __len__(self) -> int:
        return len(self.path)

    def __str__(self) -> str:
        return "/".join(self.path)

    __repr__ = __str__

    @property
    def is_root(self) -> bool:
        return not self.path

    is_empty = is_root

    @property
    def head(self) -> Optional[str]:
        return self.path[0] if self.path else None

    @property
    def next(self) -> "Path":
        return Path(self.path[1:])


class Element:
    name: Optional[str]
    offset: int

    def __init__(self, offset: int, name: Optional[str] = None):
        self.name = name
        self.offset = offset

    def __len__(self) -> int:
        raise NotImplementedError

    def __repr__(self) -> str:
        return f"[{self.offset}:{self.offset + len(self)}]"

    def token_ids(self) -> List[int]:
        raise NotImplementedError

    def position_ids(self) -> List[int]:
        raise NotImplementedError


class Parameter(Element):
    def __init__(self, offset: int, spec: ET.Element, lm: Any):
        super().__init__(offset)
        if spec.tag != "parameter":
            raise AssertionError("expected <parameter>")
        attrs = spec.attrib
        if "name" not in attrs:
            raise ValueError("Parameter name is missing")
        if "length" not in attrs:
            raise ValueError("Parameter length (in tokens) is missing")
        if not is_valid_xml_element_name(attrs["name"]):
            raise ValueError(f'Parameter name {attrs["name"]} is not valid')
        self.name = attrs["name"]
        self.length = int(attrs["length"])
        self.placeholder_token = lm.unk_token_id
        head = lm.encode(attrs["scaffold"]) if "scaffold" in attrs else []
        if len(head) > self.length:
            raise ValueError(f"Scaffold for parameter {self.name} is too long")
        self._token_ids = list(head) + [self.placeholder_token] * (self.length - len(head))
        self._position_ids = list(range(offset, offset + self.length))

    def __len__(self) -> int:
        return self.length

    def __repr__(self) -> str:
        return super().__repr__() + f" Parameter @{self.name}"

    def token_ids(self) -> List[int]:
        return self._token_ids

    def position_ids(self) -> List[int]:
        return self._position_ids


class TokenSequence(Element):
    def __init__(self, offset: int, text: Union[str, bytes], lm: Any, max_tokens: Optional[int] = None):
        super().__init__(offset)
        self.text = text
        ids = list(lm.encode(text))
        if max_tokens is not None:
            # head half + tail half, exactly as the reference slices (schema.py:167-168)
            ids = ids[:max_tokens // 2] + ids[-max_tokens // 2:]
        self._token_ids = ids
        self._position_ids = list(range(offset, offset + len(ids)))

    def __len__(self) -> int:
        return len(self._token_ids)

    def __repr__(self) -> str:
        return super().__repr__() + f" Text: {self.text!r}\n\t{self._token_ids!r}"

    def token_ids(self) -> List[int]:
        return self._token_ids

    def position_ids(self) -> List[int]:
        return self._position_ids


class UnionModule(Element):
    def __init__(self, offset: int, spec: ET.Element, lm: Any, max_tokens: Optional[int] = None):
        super().__init__(offset)
        if spec.tag != "union":
            raise AssertionError("expected <union>")
        self.modules: List[Module] = []
        for child in spec:
            if child.tag != "module":
                raise ValueError("Only <module> tags are allowed in union")
            self.modules.append(Module(offset, child, lm, max_tokens=max_tokens))
        self.length = max((len(m) for m in self.modules), default=0)
        self.scaffold_name: Optional[str] = None
        if "scaffold" in spec.attrib:
            if self.select(spec.attrib["scaffold"]) is None:
                raise ValueError(f'Union scaffold {spec.attrib["scaffold"]} is not found in union')
            self.scaffold_name = spec.attrib["scaffold"]

    def __len__(self) -> int:
        return self.length

    def __repr__(self) -> str:
        return "\n".join([super().__repr__() + " Union"] + [repr_indent(m) for m in self.modules])

    def token_ids(self) -> List[int]:
        raise ValueError("Cannot get token_ids() on union. Try again on its scaffold")

    def position_ids(self) -> List[int]:
        raise ValueError("Cannot get position_ids() on union. Try again on its scaffold")

    def select(self, path: Union[str, Path, None]) -> Optional["Module"]:
        if path is None:
            return None
        if isinstance(path, str):
            path = Path(path)
        if path.is_root:
            raise ValueError("Cannot select root of union")
        for m in self.modules:
            if m.name == path.head:
                return m if len(path) == 1 else m.select(path.next)
        return None


class Module(Element):
    def __init__(self, offset: int, spec: Union[str, ET.Element], lm: Any, is_root: bool = False,
                 max_tokens: Optional[int] = None):
        super().__init__(offset)
        self.children: List[Element] = []
        self.length = 0
        self.cache = True
        self._is_root = is_root
        self._contains_union = False
        if isinstance(spec, str):
            spec = loads_xml(spec)
        self._build(spec, lm, max_tokens)

    def _append_text(self, cursor: int, raw: Optional[str], lm: Any, max_tokens: Optional[int]) -> int:
        if raw is None:
            return cursor
        text = compact_surrounding_spaces(raw)
        if len(text) == 0:
            return cursor
        seq = TokenSequence(cursor, text, lm, max_tokens=max_tokens)
        self.children.append(seq)
        return cursor + len(seq)

    def _build(self, root: ET.Element, lm: Any, max_tokens: Optional[int]) -> None:
        expected = "schema" if self._is_root else "module"
        if root.tag != expected:
            raise AssertionError(f"expected <{expected}>, got <{root.tag}>")
        if "name" not in root.attrib:
            raise ValueError("Module name is missing")
        if not is_valid_xml_element_name(root.attrib["name"]):
            raise ValueError(f'Module name {root.attrib["name"]} is not valid')
        if not self._is_root and "cache" in root.attrib:
            self.cache = root.attrib["cache"] == "true"
        self.name = root.attrib["name"]

        cursor = self.offset
        if "src" in root.attrib:
            src = pathlib.Path(root.attrib["src"])
            if not src.exists():
                raise ValueError(f"Module source file {src} does not exist")
            cursor = self._append_text(cursor, src.read_text(), lm, max_tokens)
        cursor = self._append_text(cursor, root.text, lm, max_tokens)

        for child in root:
            tag = child.tag
            if tag == "module":
                node: Element = Module(cursor, child, lm, max_tokens=max_tokens)
                self._contains_union = self._contains_union or node._contains_union
                if node.name in [m.name for m in self.modules()]:
                    raise ValueError(f"Module {node.name} is already defined")
            elif tag == "union":
                node = UnionModule(cursor, child, lm, max_tokens=max_tokens)
                self._contains_union = True
                taken = [m.name for m in self.modules()]
                for member in node.modules:
                    if member.name in taken:
                        raise ValueError(f"Module {member.name} is already defined")
            elif tag == "parameter":
                if self._is_root:
                    raise ValueError("Parameters are not allowed in schema")
                node = Parameter(cursor, child, lm)
                if node.name in [p.name for p in self.parameters()]:
                    raise ValueError(f"Parameter {node.name} is already defined")
            else:
                # any other tag is literal text (schema.py:362-363); like lxml's tostring the serialisation
                # carries the element's tail.  The reference passes bytes to the tokenizer here (which real
                # HF tokenizers reject), we decode to str.
                node = TokenSequence(cursor, dumps_xml(child).decode("utf-8"), lm, max_tokens=max_tokens)
            self.children.append(node)
            cursor += len(node)
            cursor = self._append_text(cursor, child.tail, lm, max_tokens)
        self.length = cursor - self.offset

    def __len__(self) -> int:
        return self.length

    def __repr__(self) -> str:
        return "\n".join([super().__repr__() + f" Module @{self.name}"] + [repr_indent(c) for c in self.children])

    def token_ids(self) -> List[int]:
        if self._contains_union:
            raise ValueError("Cannot get token_ids() on module that contains union. Try again on its scaffold")
        return [t for c in self.children for t in c.token_ids()]

    def position_ids(self) -> List[int]:
        if self._contains_union:
            raise ValueError("Cannot get position_ids() on module that contains union. Try again on its scaffold")
        return [t for c in self.children for t in c.position_ids()]

    def get_scaffold(self, *paths: Path) -> "Scaffold":
        return Scaffold(self, *paths)

    def select(self, path: Union[str, Path]) -> Optional["Module"]:
        if isinstance(path, str):
            path = Path(path)
        if path.is_root:
            return self
        for m in self.modules():
            if m.name == path.head:
                return m if len(path) == 1 else m.select(path.next)
        return None

    def modules(self) -> List["Module"]:
        found: List[Module] = []
        for c in self.children:
            if type(c) is Module:
                found.append(c)
            elif type(c) is UnionModule:
                found.extend(c.modules)
        return found

    def parameters(self) -> List[Parameter]:
        return [c for c in self.children if type(c) is Parameter]

    def token_sequences(self) -> List[TokenSequence]:
        return [c for c in self.children if type(c) is TokenSequence]

    def contains_union(self) -> bool:
        return self._contains_union


class Scaffold(Element):
    """Union-free selection of a module: for every union, the member named by `paths` or its default."""

    def __init__(self, module: Module, *paths: Path):
        super().__init__(module.offset, module.name)
        self.module = module
        self.children: List[Element] = []
        for e in module.children:
            if type(e) is UnionModule:
                wanted = [p for p in paths if e.select(p.head)]
                names = list({p.head for p in wanted})
                if len(names) > 1:
                    raise ValueError("Union cannot have multiple names in scaffold")
                chosen = names[0] if wanted else e.scaffold_name
                if chosen is None:
                    continue
                self.children.append(Scaffold(e.select(chosen), *[p.next for p in wanted]))
            elif type(e) is Module:
                self.children.append(Scaffold(e, *[p.next for p in paths if p.head == e.name]))
            else:
                self.children.append(e)

    def __len__(self) -> int:
        return self.module.length

    def __repr__(self) -> str:
        return "\n".join([super().__repr__() + f" Scaffold @{self.name}"] + [repr_indent(c) for c in self.children])

    def token_ids(self) -> List[int]:
        return [t for c in self.children for t in c.token_ids()]

    def position_ids(self) -> List[int]:
        return [t for c in self.children for t in c.position_ids()]

    def select(self, path: Union[str, Path]) -> Optional["Scaffold"]:
        if isinstance(path, str):
            path = Path(path)
        if path.is_root:
            return self
        for e in self.children:
            if type(e) is Scaffold and e.name == path.head:
                return e if len(path) == 1 else e.select(path.next)
        return None

    def all_token_sequences(self) -> List[TokenSequence]:
        out: List[TokenSequence] = []
        for e in self.children:
            if type(e) is Scaffold:
                out.extend(e.all_token_sequences())
            elif type(e) is TokenSequence:
                out.append(e)
        return out


class Schema(Module):
    """Root module; parameters are not allowed at this level."""

    def __init__(self, spec: Union[str, ET.Element], lm: Any, max_tokens: Optional[int] = None):
        super().__init__(0, spec, lm, is_root=True, max_tokens=max_tokens)
        self.lm = lm
