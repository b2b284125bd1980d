"""Generates tests/golden/filter_golden.json by running THE REFERENCE's own filter pipeline in the build container:
marqo filter string -> MarqoFilterStringParser (src/marqo/core/search/search_filter.py:208-...) ->
UnstructuredVespaIndex._get_filter_term (src/marqo/core/unstructured_vespa_index/unstructured_vespa_index.py:135-226),
i.e. the exact YQL text that follows `... nearestNeighbor(...)) AND ` in a filtered tensor query (:62-66).

    python tests/golden/make_filter_golden.py
"""
import json
import sys
from pathlib import Path
from types import SimpleNamespace

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import _reference_import  # noqa: E402

_reference_import.install()
import torchaudio  # noqa: E402

if not hasattr(torchaudio, "set_audio_backend"):
    torchaudio.set_audio_backend = lambda *a, **k: None

from marqo.core.search.search_filter import MarqoFilterStringParser  # noqa: E402
from marqo.core.unstructured_vespa_index.unstructured_vespa_index import UnstructuredVespaIndex  # noqa: E402

FILTERS = [
    "color:red",
    "color:(dark red)",
    "price:[10 TO 20]",
    "price:[10.5 TO *]",
    "price:[* TO 3]",
    "in_stock:true",
    "in_stock:false AND color:red",
    "color:red OR color:blue",
    "NOT color:red",
    "(color:red OR color:blue) AND price:[0 TO 100]",
    "NOT (color:red AND in_stock:true)",
    "tags:sale",
    "year:2024",
    "rating:4.5",
    "_id:doc7",
    "meta.size:large",
    'title:(say \\"hi\\")',
    "a:1 AND (b:2 OR (c:3 AND NOT d:4))",
    "color:RED",
]
out = []
parser = MarqoFilterStringParser()
for f in FILTERS:
    tree = parser.parse(f)
    q = SimpleNamespace(filter=tree)
    out.append({"filter": f, "yql": UnstructuredVespaIndex._get_filter_term(q)})
(HERE / "filter_golden.json").write_text(json.dumps(out, indent=1))
for o in out[:6]:
    print(o)
print(len(out))
