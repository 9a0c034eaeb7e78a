"""``paragraphs.bin`` / ``paragraphs.pos`` of the reference's data store v2 (host side, no device).

nidx/nidx_vector/src/data_store/v2/paragraph_store.rs:31-41,74-166: one ``StoredParagraph {key: &str, labels: Vec<&str>,
metadata: &[u8], first_vector: u32, num_vectors: u32}`` per paragraph, serialised with wincode configured to match
``bincode::config::standard()`` of bincode 2.x (utils.rs:25-28): little endian, variable-length integers, lengths as u64
varints.  ``paragraphs.pos`` holds the u32 LE start offset of every record.

bincode 2 varint (published format; wincode itself is not vendored in the reference tree, so byte parity with a
reference-written file is unpinned -- the tests pin this module to the published encoding):
  u < 251 -> 1 byte;  < 2^16 -> 251, u16 LE;  < 2^32 -> 252, u32 LE;  < 2^64 -> 253, u64 LE.
"""
from __future__ import annotations

import os
import struct
from typing import Iterable, List, Optional, Sequence, Tuple

FILENAME_DATA = "paragraphs.bin"   # paragraph_store.rs:31
FILENAME_POS = "paragraphs.pos"    # paragraph_store.rs:32

# (key, labels, metadata, first_vector, num_vectors)
Paragraph = Tuple[str, Sequence[str], Optional[bytes], int, int]


def encode_varint(u: int) -> bytes:
    if u < 0:
        raise ValueError("unsigned only")
    if u < 251:
        return bytes([u])
    if u < 1 << 16:
        return b"\xfb" + struct.pack("<H", u)
    if u < 1 << 32:
        return b"\xfc" + struct.pack("<I", u)
    if u < 1 << 64:
        return b"\xfd" + struct.pack("<Q", u)
    raise ValueError("u128 is not used by this format")


def decode_varint(buf, pos: int) -> Tuple[int, int]:
    """-> (value, next position)."""
    tag = buf[pos]
    if tag < 251:
        return tag, pos + 1
    if tag == 251:
        return struct.unpack_from("<H", buf, pos + 1)[0], pos + 3
    if tag == 252:
        return struct.unpack_from("<I", buf, pos + 1)[0], pos + 5
    if tag == 253:
        return struct.unpack_from("<Q", buf, pos + 1)[0], pos + 9
    raise ValueError(f"unsupported varint tag {tag} at {pos}")


def _bytes_field(b: bytes) -> bytes:
    return encode_varint(len(b)) + b


def encode_paragraph(key: str, labels: Sequence[str], metadata: Optional[bytes], first_vector: int, num_vectors: int) -> bytes:
    """StoredParagraph in field order (paragraph_store.rs:34-41; metadata None is stored as the empty slice, :57)."""
    out = bytearray(_bytes_field(key.encode("utf-8")))
    out += encode_varint(len(labels))
    for label in labels:
        out += _bytes_field(label.encode("utf-8"))
    out += _bytes_field(metadata or b"")
    out += encode_varint(first_vector) + encode_varint(num_vectors)
    return bytes(out)


def decode_paragraph(buf, pos: int = 0) -> Tuple[Paragraph, int]:
    n, pos = decode_varint(buf, pos)
    key = bytes(buf[pos:pos + n]).decode("utf-8")
    pos += n
    n_labels, pos = decode_varint(buf, pos)
    labels = []
    for _ in range(n_labels):
        n, pos = decode_varint(buf, pos)
        labels.append(bytes(buf[pos:pos + n]).decode("utf-8"))
        pos += n
    n, pos = decode_varint(buf, pos)
    metadata = bytes(buf[pos:pos + n])
    pos += n
    first_vector, pos = decode_varint(buf, pos)
    num_vectors, pos = decode_varint(buf, pos)
    return (key, labels, metadata or None, first_vector, num_vectors), pos


def write_paragraphs(directory: str, paragraphs: Iterable[Paragraph]) -> int:
    """ParagraphStoreWriter (paragraph_store.rs:112-166): records back to back, one u32 LE start per record.  -> count."""
    data, pos, count = bytearray(), bytearray(), 0
    for key, labels, metadata, first_vector, num_vectors in paragraphs:
        if len(data) >= 1 << 32:
            raise ValueError("paragraphs.bin exceeds the u32 offsets of paragraphs.pos")
        pos += struct.pack("<I", len(data))
        data += encode_paragraph(key, labels, metadata, first_vector, num_vectors)
        count += 1
    with open(os.path.join(directory, FILENAME_DATA), "wb") as f:
        f.write(data)
    with open(os.path.join(directory, FILENAME_POS), "wb") as f:
        f.write(pos)
    return count


def read_paragraphs(directory: str) -> List[Paragraph]:
    """ParagraphStore::get_paragraph for every address (paragraph_store.rs:97-103, stored_elements :109)."""
    with open(os.path.join(directory, FILENAME_POS), "rb") as f:
        pos = f.read()
    with open(os.path.join(directory, FILENAME_DATA), "rb") as f:
        data = f.read()
    if len(pos) % 4:
        raise ValueError("paragraphs.pos is not a whole number of u32 offsets")
    out = []
    for (start,) in struct.iter_unpack("<I", pos):
        if start > len(data):
            raise ValueError("paragraphs.pos points past the end of paragraphs.bin")
        try:
            paragraph, _ = decode_paragraph(data, start)
        except (IndexError, struct.error, UnicodeDecodeError) as e:
            raise ValueError(f"corrupt paragraph record at {start}: {e}") from None
        out.append(paragraph)
    return out
