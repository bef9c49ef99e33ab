"""Reader of the DGPUDMP1 interchange file (written by oracle/ref_driver export from a reference index, or by
dgpu_write_synthetic_dump). Pure data plumbing for tests and tools: numpy views over the file."""
import gzip
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np


@dataclass
class FieldSegment:
    has_terms: bool = False
    norms: Optional[np.ndarray] = None          # int8[max_doc]
    sum_total_term_freq: int = -1
    sum_doc_freq: int = -1
    doc_count: int = 0
    terms: Dict[bytes, tuple] = field(default_factory=dict)  # term -> (docs int32[], freqs int32[], ttf)


@dataclass
class Segment:
    max_doc: int
    doc_base: int
    fields: Dict[str, FieldSegment]
    dv: Dict[str, np.ndarray]


@dataclass
class Dump:
    field_names: List[str]
    dv_names: List[str]
    segments: List[Segment]

    @property
    def max_doc(self):
        return sum(s.max_doc for s in self.segments)


def read_dump(path) -> Dump:
    path = str(path)
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        buf = f.read()
    if buf[:8] != b"DGPUDMP1":
        raise ValueError("not a DGPUDMP1 file")
    pos = 8

    def u32():
        nonlocal pos
        v = struct.unpack_from("<I", buf, pos)[0]
        pos += 4
        return v

    def i64():
        nonlocal pos
        v = struct.unpack_from("<q", buf, pos)[0]
        pos += 8
        return v

    def u8():
        nonlocal pos
        v = buf[pos]
        pos += 1
        return v

    def string():
        nonlocal pos
        n = u32()
        s = buf[pos:pos + n]
        pos += n
        return s

    n_seg, n_fields = u32(), u32()
    fields = [string().decode() for _ in range(n_fields)]
    n_dv = u32()
    dvs = [string().decode() for _ in range(n_dv)]
    segments = []
    for _ in range(n_seg):
        max_doc, doc_base = u32(), u32()
        seg_fields = {}
        for name in fields:
            fs = FieldSegment()
            fs.has_terms = bool(u8())
            if u8():
                fs.norms = np.frombuffer(buf, dtype=np.int8, count=max_doc, offset=pos).copy()
                pos += max_doc
            if fs.has_terms:
                fs.sum_total_term_freq, fs.sum_doc_freq = i64(), i64()
                fs.doc_count = struct.unpack_from("<i", buf, pos)[0]
                pos += 4
                n_terms = struct.unpack_from("<Q", buf, pos)[0]
                pos += 8
                for _t in range(n_terms):
                    term = string()
                    df = u32()
                    ttf = i64()
                    pairs = np.frombuffer(buf, dtype=np.uint32, count=2 * df, offset=pos).reshape(df, 2)
                    pos += 8 * df
                    fs.terms[term] = (pairs[:, 0].astype(np.int32), pairs[:, 1].astype(np.int32), ttf)
            seg_fields[name] = fs
        seg_dv = {}
        for name in dvs:
            if u8():
                seg_dv[name] = np.frombuffer(buf, dtype=np.int64, count=max_doc, offset=pos).copy()
                pos += 8 * max_doc
        segments.append(Segment(max_doc, doc_base, seg_fields, seg_dv))
    return Dump(fields, dvs, segments)


def build_reader_from_dump(dump: Dump, device: int = 0, local_segments=None):
    """Feeds a Dump through the dgpu_builder_* C ABI (the path a reader-side integration would use)."""
    from .api import IndexBuilder

    b = IndexBuilder()
    for si, seg in enumerate(dump.segments):
        local = local_segments is None or si in local_segments
        s = b.add_segment(seg.max_doc, seg.doc_base, local)
        for name, fs in seg.fields.items():
            if not fs.has_terms:
                continue
            b.set_field_stats(s, name, fs.sum_total_term_freq, fs.sum_doc_freq, fs.doc_count, fs.norms)
            for term, (docs, freqs, ttf) in fs.terms.items():
                if local:
                    b.add_term(s, name, term, docs, freqs, ttf=ttf)
                else:
                    b.add_term(s, name, term, None, None, doc_freq=len(docs), ttf=ttf)
        if local:
            for name, vals in seg.dv.items():
                b.add_numeric_doc_values(s, name, vals)
    return b.finish(device)
