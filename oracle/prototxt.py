"""oracle/prototxt.py -- TEST INFRASTRUCTURE (parity checker side), not product code.

Minimal protobuf text-format reader, enough for caffe_3d net definitions
(caffe_3d/src/caffe/proto/caffe.proto; read by ReadProtoFromTextFile,
caffe_3d/src/caffe/util/io.cpp).  Every field is kept as a list (protobuf
`repeated` semantics); scalar access goes through `Msg.get(name, default)`.

The product has its own C++ parser (csrc/prototxt.cpp); the two are written
independently and the tests compare what they produce.
"""
from __future__ import annotations

import re

_TOKEN = re.compile(
    r"""\s*(?:
        (?P<comment>\#[^\n]*) |
        (?P<str>"(?:[^"\\]|\\.)*"|'(?:[^'\\]|\\.)*') |
        (?P<punct>[{}\[\]:,;<>]) |
        (?P<atom>[^\s{}\[\]:,;<>"'#]+)
    )""",
    re.X,
)


class Msg(dict):
    """A parsed message: field name -> list of values (str/int/float/bool/Msg)."""

    def get1(self, name, default=None):
        v = dict.get(self, name)
        return v[-1] if v else default

    def getall(self, name):
        return dict.get(self, name, [])

    def has(self, name):
        return bool(dict.get(self, name))


def _tokens(text):
    pos = 0
    n = len(text)
    while pos < n:
        m = _TOKEN.match(text, pos)
        if not m:
            if text[pos:].strip() == "":
                return
            raise ValueError("prototxt: cannot tokenise at offset %d: %r" % (pos, text[pos:pos + 30]))
        pos = m.end()
        if m.group("comment") is not None:
            continue
        if m.group("str") is not None:
            s = m.group("str")[1:-1]
            yield ("str", bytes(s, "utf-8").decode("unicode_escape"))
        elif m.group("punct") is not None:
            yield ("punct", m.group("punct"))
        elif m.group("atom") is not None:
            yield ("atom", m.group("atom"))


def _scalar(kind, tok):
    if kind == "str":
        return tok
    if tok in ("true", "True"):
        return True
    if tok in ("false", "False"):
        return False
    try:
        if re.fullmatch(r"[-+]?\d+", tok):
            return int(tok)
        if re.fullmatch(r"[-+]?0[xX][0-9a-fA-F]+", tok):
            return int(tok, 16)
        return float(tok.rstrip("fF"))
    except ValueError:
        return tok  # enum identifier (MAX, AVE, TEST, ...)


def parse(text: str) -> Msg:
    toks = list(_tokens(text))
    i = 0

    def parse_msg(close):
        nonlocal i
        msg = Msg()
        while i < len(toks):
            kind, tok = toks[i]
            if kind == "punct" and tok == close:
                i += 1
                return msg
            if kind == "punct" and tok in ",;":
                i += 1
                continue
            if kind != "atom":
                raise ValueError("prototxt: expected field name, got %r" % (tok,))
            name = tok
            i += 1
            kind, tok = toks[i]
            if kind == "punct" and tok == ":":
                i += 1
                kind, tok = toks[i]
            if kind == "punct" and tok in "{<":
                i += 1
                msg.setdefault(name, []).append(parse_msg("}" if tok == "{" else ">"))
            elif kind == "punct" and tok == "[":
                i += 1
                while True:
                    kind, tok = toks[i]
                    if kind == "punct" and tok == "]":
                        i += 1
                        break
                    if kind == "punct" and tok == ",":
                        i += 1
                        continue
                    if kind == "punct" and tok in "{<":
                        i += 1
                        msg.setdefault(name, []).append(parse_msg("}" if tok == "{" else ">"))
                    else:
                        msg.setdefault(name, []).append(_scalar(kind, tok))
                        i += 1
            else:
                val = _scalar(kind, tok)
                i += 1
                # adjacent string literals concatenate in text format
                while kind == "str" and i < len(toks) and toks[i][0] == "str":
                    val += toks[i][1]
                    i += 1
                msg.setdefault(name, []).append(val)
        if close is not None:
            raise ValueError("prototxt: unterminated message")
        return msg

    return parse_msg(None)


def parse_file(path: str) -> Msg:
    with open(path, "r") as f:
        return parse(f.read())
