"""Minimal HOCON-subset reader for the AppearanceGen ``confs/**/*.conf`` files (pyhocon is not a dependency).

Covers what the 180 shipped confs use: nested ``name { ... }`` blocks, ``key = value``, ``#`` comments,
lists (also multi-line), trailing commas after values (``D = 4,``), unquoted strings with spaces, ``True/False``,
ints / floats (``5e-4``), and the accessor API the reference calls on pyhocon's ConfigTree (main.py:39-127):
``conf['a.b']``, ``get_int/float/bool/string(key, default=...)``, subtree-as-kwargs (``**conf['model.sdf_network']``).
"""
from __future__ import annotations

import re
from collections import OrderedDict
from typing import Any


class ConfigMissingException(KeyError):
    pass


_NO_DEFAULT = object()


class ConfigTree(OrderedDict):
    def _lookup(self, key: str):
        node: Any = self
        for part in key.split("."):
            if not isinstance(node, dict) or part not in node:
                raise ConfigMissingException(key)
            node = OrderedDict.__getitem__(node, part)
        return node

    def __getitem__(self, key):
        if isinstance(key, str) and "." in key:
            return self._lookup(key)
        try:
            return OrderedDict.__getitem__(self, key)
        except KeyError:
            raise ConfigMissingException(key)

    def get(self, key, default=_NO_DEFAULT):
        try:
            return self._lookup(key) if isinstance(key, str) else OrderedDict.__getitem__(self, key)
        except (ConfigMissingException, KeyError):
            if default is _NO_DEFAULT:
                raise ConfigMissingException(key)
            return default

    def _typed(self, key, default, conv):
        v = self.get(key, default)
        return v if v is default and default is not _NO_DEFAULT else conv(v)

    def get_int(self, key, default=_NO_DEFAULT):
        return self._typed(key, default, int)

    def get_float(self, key, default=_NO_DEFAULT):
        return self._typed(key, default, float)

    def get_string(self, key, default=_NO_DEFAULT):
        return self._typed(key, default, str)

    def get_bool(self, key, default=_NO_DEFAULT):
        def conv(v):
            if isinstance(v, bool):
                return v
            if isinstance(v, str):
                if v.lower() in ("true", "yes", "on"):
                    return True
                if v.lower() in ("false", "no", "off"):
                    return False
            raise ValueError(f"{key}: not a boolean: {v!r}")
        return self._typed(key, default, conv)

    def get_list(self, key, default=_NO_DEFAULT):
        return self._typed(key, default, list)


def _scalar(tok: str):
    t = tok.strip()
    if len(t) >= 2 and t[0] == t[-1] and t[0] in "\"'":
        return t[1:-1]
    if t in ("True", "true"):
        return True
    if t in ("False", "false"):
        return False
    if t in ("null", "None"):
        return None
    try:
        return int(t)
    except ValueError:
        pass
    try:
        return float(t)
    except ValueError:
        return t


def _strip_comment(line: str) -> str:
    out, q = [], None
    i = 0
    while i < len(line):
        ch = line[i]
        if q:
            if ch == q:
                q = None
        elif ch in "\"'":
            q = ch
        elif ch == "#" or line.startswith("//", i):
            break
        out.append(ch)
        i += 1
    return "".join(out)


def parse_string(text: str) -> ConfigTree:
    text = "\n".join(_strip_comment(l) for l in text.splitlines())
    pos = 0
    n = len(text)

    def skip_ws():
        nonlocal pos
        while pos < n and text[pos] in " \t\r\n,":
            pos += 1

    def parse_list():
        nonlocal pos
        assert text[pos] == "["
        pos += 1
        items = []
        while True:
            skip_ws()
            if pos >= n:
                raise ValueError("unterminated list")
            if text[pos] == "]":
                pos += 1
                return items
            if text[pos] == "[":
                items.append(parse_list())
            elif text[pos] == "{":
                pos += 1
                items.append(parse_object(True))
            else:
                m = re.compile(r"[^,\]\n]+").match(text, pos)
                items.append(_scalar(m.group(0)))
                pos = m.end()

    def parse_value():
        nonlocal pos
        while pos < n and text[pos] in " \t":
            pos += 1
        if pos < n and text[pos] == "[":
            return parse_list()
        if pos < n and text[pos] == "{":
            pos += 1
            return parse_object(True)
        # an unquoted value runs to the end of the line; braces inside it are kept when balanced (the colab conf has
        # prompts like "a 3D rendering of a {TOREPLACE} in unreal engine"), an unbalanced '}' closes the enclosing block
        start, depth = pos, 0
        next_key = re.compile(r",\s*[A-Za-z0-9_.\-\"']+\s*[={:]")
        while pos < n and text[pos] != "\n":
            if text[pos] == "," and depth == 0 and next_key.match(text, pos):
                break                      # "b = 1, c = 2" on one line: the comma separates two assignments
            if text[pos] == "{":
                depth += 1
            elif text[pos] == "}":
                if depth == 0:
                    break
                depth -= 1
            pos += 1
        raw = text[start:pos]
        return _scalar(raw.rstrip().rstrip(",").rstrip())

    def parse_object(nested: bool) -> ConfigTree:
        nonlocal pos
        tree = ConfigTree()
        while True:
            skip_ws()
            if pos >= n:
                if nested:
                    raise ValueError("unterminated block")
                return tree
            if text[pos] == "}":
                pos += 1
                return tree
            m = re.compile(r"[A-Za-z0-9_.\-\"']+").match(text, pos)
            if not m:
                raise ValueError(f"unexpected character {text[pos]!r} at offset {pos}")
            key = m.group(0).strip("\"'")
            pos = m.end()
            while pos < n and text[pos] in " \t":
                pos += 1
            if pos < n and text[pos] == "{":
                pos += 1
                val = parse_object(True)
            elif pos < n and text[pos] in "=:":
                pos += 1
                val = parse_value()
            else:
                raise ValueError(f"expected '=' or '{{' after key {key!r}")
            node = tree
            parts = key.split(".")
            for part in parts[:-1]:
                node = node.setdefault(part, ConfigTree())
            if isinstance(val, dict) and isinstance(node.get(parts[-1], None), dict):
                node[parts[-1]].update(val)
            else:
                node[parts[-1]] = val

    return parse_object(False)


def parse_file(path: str) -> ConfigTree:
    with open(path) as f:
        return parse_string(f.read())
