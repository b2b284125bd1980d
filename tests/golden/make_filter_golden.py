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
    # Recommender._get_exclusion_filter (core/search/recommender.py:205-214), alone and on top of a user filter
    "NOT (_id:(d0) OR _id:(d1) OR _id:(d2))",
    "(color:red) AND NOT (_id:(d0) OR _id:(doc7))",
]
out = []
parser = MarqoFilterStringParser()
for f in FILTERS:
    tree = parser.parse(f)
    q = SimpleNamespace(filter=tree)
    out.append({"filter": f, "yql": UnstructuredVespaIndex._get_filter_term(q)})
# ---- structured indexes: StructuredVespaIndex._get_filter_term (structured_vespa_index.py:690-793) over an index whose
# filterable fields are color / title (text), price / rating (float), year (int), in_stock (bool), tags (array<text>)
import time  # noqa: E402

from marqo.core.models.marqo_index import (AudioPreProcessing, DistanceMetric, Field, FieldFeature, FieldType,  # noqa: E402
                                           HnswConfig, ImagePreProcessing, Model, StructuredMarqoIndex, TensorField,
                                           TextPreProcessing, TextSplitMethod, VectorNumericType, VideoPreProcessing)
from marqo.core.structured_vespa_index.structured_vespa_index import StructuredVespaIndex  # noqa: E402


def _field(name, typ):
    return Field(name=name, type=typ, features=[FieldFeature.Filter], filter_field_name=f"marqo__filter_{name}",
                 lexical_field_name=None)


index = StructuredMarqoIndex(
    name="i", schema_name="s", model=Model(name="hf/all_datasets_v4_MiniLM-L6"), normalize_embeddings=True,
    text_preprocessing=TextPreProcessing(split_length=2, split_overlap=0, split_method=TextSplitMethod.Sentence),
    image_preprocessing=ImagePreProcessing(patch_method=None),
    video_preprocessing=VideoPreProcessing(split_length=20, split_overlap=1),
    audio_preprocessing=AudioPreProcessing(split_length=20, split_overlap=1),
    distance_metric=DistanceMetric.PrenormalizedAngular, vector_numeric_type=VectorNumericType.Float,
    hnsw_config=HnswConfig(ef_construction=128, m=16),
    fields=[_field("color", FieldType.Text), _field("title", FieldType.Text), _field("price", FieldType.Float),
            _field("rating", FieldType.Float), _field("year", FieldType.Int), _field("in_stock", FieldType.Bool),
            _field("tags", FieldType.ArrayText)],
    tensor_fields=[TensorField(name="title", chunk_field_name="marqo__chunks_title",
                               embeddings_field_name="marqo__embeddings_title")],
    marqo_version="2.12.0", created_at=time.time(), updated_at=time.time(), version=None)
svi = StructuredVespaIndex(index)
STRUCTURED = ["color:red", "price:[10 TO 20]", "price:[10.5 TO *]", "price:[* TO 3]", "in_stock:true",
              "in_stock:false AND color:red", "color:red OR color:blue", "NOT color:red",
              "(color:red OR color:blue) AND price:[0 TO 100]", "NOT (color:red AND in_stock:true)", "tags:sale",
              "year:2024", "rating:4.5", "_id:doc7", 'title:(say \\"hi\\")', "color in (red, blue)",
              "year in (2023, 2024)", "color:RED", "NOT _id in (d0, d1)"]
structured = []
for f in STRUCTURED:
    structured.append({"filter": f, "yql": svi._get_filter_term(SimpleNamespace(filter=parser.parse(f)))})
(HERE / "filter_golden_structured.json").write_text(json.dumps(structured, indent=1))
print(structured[-1], len(structured))
(HERE / "filter_golden.json").write_text(json.dumps(out, indent=1))
for o in out[:6]:
    print(o)
print(len(out))
