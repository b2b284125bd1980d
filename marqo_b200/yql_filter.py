"""Host-side evaluation of the YQL filter text Marqo appends to a tensor query (` AND <filter>`,
src/marqo/core/unstructured_vespa_index/unstructured_vespa_index.py:62-66), for unstructured and semi-structured indexes
(the latter reuses the same generator: semi_structured_vespa_index.py:67-69).

Grammar — exactly what `UnstructuredVespaIndex._get_filter_term` (:135-226) and, for structured indexes,
`StructuredVespaIndex._get_filter_term` (src/marqo/core/structured_vespa_index/structured_vespa_index.py:690-793) emit:

    expr  := '(' expr {('AND' | 'OR') expr} ')'  |  '!(' expr ')'  |  atom
    atom  := FIELD 'contains' STRING                                       marqo__id, marqo__string_array, marqo__filter_<f>
           | FIELD 'contains' 'sameElement(' 'key' 'contains' STRING ',' cond {',' cond} ')'       unstructured map fields
           | FIELD ('>=' | '<=') NUMBER                                                              structured ranges
           | FIELD 'in' '(' (STRING | NUMBER) {',' (STRING | NUMBER)} ')'                            structured IN
    cond  := 'value' 'contains' STRING  |  'value' ('=' | '>=' | '<=') NUMBER

Semantics are Vespa's for the schemas the reference generates (unstructured_vespa_schema.py:86-140,
structured_vespa_schema.py:93-100): every filtered field is an attribute without `match: cased`, so string matching is
whole-value and case-insensitive; `contains` on a numeric attribute is equality; a multi-value field matches when any
element does; `sameElement` requires key and value conditions to hold for the SAME map entry.  Anything outside this
grammar raises FilterSyntaxError and the adapter delegates the query instead of answering it.
"""
from __future__ import annotations

import re
from typing import Any, Callable, Dict, List, Tuple


class FilterSyntaxError(ValueError):
    pass


_TOKEN = re.compile(r'\s*(?:("(?:[^"\\]|\\.)*")|(!\(|\(|\)|,|>=|<=|=)|([A-Za-z_][A-Za-z0-9_.]*)|(-?\d+(?:\.\d+)?(?:[eE][-+]?\d+)?))')

# unstructured / semi-structured filter attributes (unstructured_vespa_index/common.py:3-13); structured ones are
# `marqo__filter_<field>` (structured_vespa_schema.py:93-100)
_FILTER_FIELDS = {"marqo__id", "marqo__short_string_fields", "marqo__string_array", "marqo__int_fields",
                  "marqo__float_fields", "marqo__bool_fields"}

Doc = Dict[str, Any]
Pred = Callable[[Doc], bool]


def _tokens(text: str) -> List[Tuple[str, str]]:
    out, pos = [], 0
    text = text.strip()
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m or m.end() == pos:
            raise FilterSyntaxError(f"unexpected text at {pos}: {text[pos:pos + 30]!r}")
        if m.group(1) is not None:
            out.append(("str", re.sub(r"\\(.)", r"\1", m.group(1)[1:-1])))     # undo escape(): \\ -> \, \" -> "
        elif m.group(2) is not None:
            out.append(("sym", m.group(2)))
        elif m.group(3) is not None:
            out.append(("word", m.group(3)))
        else:
            out.append(("num", m.group(4)))
        pos = m.end()
    return out


def _fold(s: Any) -> str:
    return str(s).lower()


class _Parser:
    def __init__(self, toks: List[Tuple[str, str]]):
        self.t, self.i = toks, 0

    def peek(self, k: int = 0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def take(self, kind: str = None, val: str = None):
        tok = self.peek()
        if (kind and tok[0] != kind) or (val is not None and tok[1] != val):
            raise FilterSyntaxError(f"expected {val or kind}, found {tok[1]!r}")
        self.i += 1
        return tok[1]

    def expr(self) -> Pred:
        if self.peek() == ("sym", "!("):
            self.take()
            inner = self.expr()
            self.take("sym", ")")
            return lambda d: not inner(d)
        if self.peek() != ("sym", "("):
            return self.atom()
        self.take("sym", "(")
        terms, ops = [self.expr()], []
        while self.peek()[0] == "word" and self.peek()[1] in ("AND", "OR"):
            ops.append(self.take())
            terms.append(self.expr())
        self.take("sym", ")")
        if len(set(ops)) > 1:   # the generators never mix operators inside one pair of parentheses
            raise FilterSyntaxError("mixed AND / OR without parentheses")
        if not ops:
            return terms[0]
        if ops[0] == "AND":
            return lambda d: all(t(d) for t in terms)
        return lambda d: any(t(d) for t in terms)

    def atom(self) -> Pred:
        field = self.take("word")
        if field not in _FILTER_FIELDS and not field.startswith("marqo__filter_"):
            # e.g. `default contains "x"` is a LEXICAL term on the bm25 fieldset, not a filter
            raise FilterSyntaxError(f"{field!r} is not a filter attribute")
        if self.peek()[0] == "sym" and self.peek()[1] in (">=", "<="):
            op, num = self.take(), float(self.take("num"))
            return lambda d: any((_num(v) >= num) if op == ">=" else (_num(v) <= num) for v in _values(d.get(field)))
        if self.peek() == ("word", "in"):
            self.take()
            self.take("sym", "(")
            wanted = [self._literal()]
            while self.peek() == ("sym", ","):
                self.take()
                wanted.append(self._literal())
            self.take("sym", ")")
            return lambda d: any(_equal(v, w) for v in _values(d.get(field)) for w in wanted)
        self.take("word", "contains")
        if self.peek()[0] == "str":
            want = self.take("str")
            return lambda d: any(_equal(v, want) for v in _values(d.get(field)))
        self.take("word", "sameElement")
        self.take("sym", "(")
        self.take("word", "key")
        self.take("word", "contains")
        key = _fold(self.take("str"))
        conds: List[Callable[[Any], bool]] = []
        while self.peek() == ("sym", ","):
            self.take()
            self.take("word", "value")
            if self.peek() == ("word", "contains"):
                self.take()
                s = _fold(self.take("str"))
                conds.append(lambda v, s=s: _fold(v) == s)
            else:
                op = self.take("sym")
                num = float(self.take("num"))
                if op == "=":
                    conds.append(lambda v, n=num: _num(v) == n)
                elif op == ">=":
                    conds.append(lambda v, n=num: _num(v) >= n)
                elif op == "<=":
                    conds.append(lambda v, n=num: _num(v) <= n)
                else:
                    raise FilterSyntaxError(f"unsupported comparison {op!r}")
        self.take("sym", ")")
        if not conds:
            raise FilterSyntaxError("sameElement without a value condition")

        def same_element(d: Doc) -> bool:
            m = d.get(field)
            if not isinstance(m, dict):
                return False
            return any(_fold(k) == key and all(c(v) for c in conds) for k, v in m.items())
        return same_element

    def _literal(self):
        kind, val = self.peek()
        if kind == "str":
            return self.take()
        if kind == "num":
            return float(self.take())
        raise FilterSyntaxError(f"expected a string or a number, found {val!r}")


def _values(v: Any) -> List[Any]:
    if v is None:
        return []
    return list(v) if isinstance(v, (list, tuple)) else [v]


def _equal(stored: Any, wanted: Any) -> bool:
    """`contains` / `in` on an attribute: numeric fields compare as numbers, strings whole-value and uncased."""
    if isinstance(stored, bool):
        stored = int(stored)
    if isinstance(stored, (int, float)):
        return _num(wanted) == float(stored)
    return not isinstance(wanted, float) and _fold(stored) == _fold(wanted)


def _num(v: Any) -> float:
    try:
        return float(v)
    except (TypeError, ValueError):
        return float("nan")


def compile_filter(yql_filter: str) -> Pred:
    """YQL filter text -> predicate over a document's stored Vespa fields.  Raises FilterSyntaxError."""
    p = _Parser(_tokens(yql_filter))
    pred = p.expr()
    if p.peek()[0] != "eof":
        raise FilterSyntaxError(f"trailing text: {p.peek()[1]!r}")
    return pred
