"""Host-side `dna`, `kmer` and `qkmer` values with the reference's text I/O rules.

These mirror the scalar input functions of the extension -- `dna_in`/`dna_make`
(dna.c:178-202, 220), `kmer_in`/`kmer_make` (dna.c:487-515, 528) and
`qkmer_in`/`qkmer_make` (dna.c:908-930, 932) -- which stay host code in the
reference too ("scalar glue").  They build inputs for the GPU
path; none of the hot path runs here.
"""
import numpy as np

_CODE = np.full(256, 255, dtype=np.uint8)
for _c, _v in (("A", 0), ("T", 1), ("C", 2), ("G", 3)):
    _CODE[ord(_c)] = _v
_LETTERS = np.frombuffer(b"ATCG", dtype=np.uint8)
QKMER_ALPHABET = "ATCGUWSMKRYBDHVN"  # dna.c:889-891


class DnaError(ValueError):
    """An `ereport(ERROR, ...)` of the reference, or a libdnagpu status code."""

    def __init__(self, message, code=None):
        super().__init__(message)
        self.code = code


def pack_bases(codes: np.ndarray) -> np.ndarray:
    """2-bit codes (uint8, 0..3) -> little-endian packed uint64 words (dna.c:116-123)."""
    n = codes.size
    n_words = (n + 31) // 32
    padded = np.zeros(n_words * 32, dtype=np.uint8)
    padded[:n] = codes
    b = padded.reshape(-1, 4)
    by = (b[:, 0] | (b[:, 1] << 2) | (b[:, 2] << 4) | (b[:, 3] << 6)).astype(np.uint8)
    return by.view("<u8").copy()


def unpack_bases(words: np.ndarray, n: int) -> np.ndarray:
    by = np.ascontiguousarray(words, dtype="<u8").view(np.uint8)
    codes = np.empty(by.size * 4, dtype=np.uint8)
    for j in range(4):
        codes[j::4] = (by >> (2 * j)) & 3
    return codes[:n]


class Dna:
    """A `dna` value: `length` bases packed 32 per uint64 word (struct Dna, dna.c:42-47)."""

    __slots__ = ("words", "length")

    def __init__(self, text):
        if isinstance(text, Dna):
            self.words, self.length = text.words, text.length
            return
        raw = np.frombuffer(text.encode("ascii", "replace") if isinstance(text, str) else bytes(text),
                            dtype=np.uint8)
        if raw.size == 0:
            raise DnaError("DNA sequence cannot be empty")  # dna.c:160-161
        codes = _CODE[raw]
        bad = np.nonzero(codes == 255)[0]
        if bad.size:
            raise DnaError(f"Invalid character in DNA sequence: {chr(raw[bad[0]])}")  # dna.c:166
        self.words = pack_bases(codes)
        self.length = int(raw.size)

    @classmethod
    def from_words(cls, words, length):
        self = cls.__new__(cls)
        self.words = np.ascontiguousarray(words, dtype=np.uint64)
        self.length = int(length)
        if self.words.size < (self.length + 31) // 32:
            raise DnaError("too few words for the stated length")
        return self

    def __len__(self):
        return self.length

    def __str__(self):  # dna_out -> decode_dna, dna.c:135-152
        return _LETTERS[unpack_bases(self.words, self.length)].tobytes().decode()

    def __eq__(self, other):  # dna_eq_internal, dna.c:334-350
        return (isinstance(other, Dna) and self.length == other.length and
                np.array_equal(self.words[:(self.length + 31) // 32],
                               other.words[:(other.length + 31) // 32]))

    def __repr__(self):
        s = str(self)
        return f"Dna({s[:40]!r}{'...' if len(s) > 40 else ''}, length={self.length})"


class Kmer:
    """A `kmer` value: (length, bit_sequence) (struct Kmer, dna.c:61-65)."""

    __slots__ = ("bits", "length")

    def __init__(self, text=None, *, bits=None, length=None):
        if text is None:
            if not 1 <= int(length) <= 32:
                raise DnaError("K-mer length must be between 1 and 32 nucleotides")  # dna.c:401-402
            self.bits, self.length = int(bits), int(length)
            return
        if isinstance(text, Kmer):
            self.bits, self.length = text.bits, text.length
            return
        if text == "":
            raise DnaError("K-mer sequence cannot be empty")  # dna.c:460-461
        if len(text) > 32:
            raise DnaError("K-mer length cannot exceed 32 nucleotides")  # dna.c:466-467
        bits = 0
        for i, ch in enumerate(text):  # encode_kmer, dna.c:405-417 ('X' -> 00, dna.c:413)
            if ch not in "ATCGX":
                raise DnaError(f"Invalid character in K-mer sequence: '{ch}'")  # dna.c:473
            bits |= {"A": 0, "T": 1, "C": 2, "G": 3, "X": 0}[ch] << (2 * i)
        self.bits, self.length = bits, len(text)

    def __str__(self):  # decode_kmer, dna.c:428-452
        return "".join("ATCG"[(self.bits >> (2 * i)) & 3] for i in range(self.length))

    def __eq__(self, other):  # kmer_eq_internal, dna.c:655-668
        return isinstance(other, Kmer) and self.length == other.length and self.bits == other.bits

    def __hash__(self):
        return hash((self.length, self.bits))

    def __repr__(self):
        return f"Kmer({str(self)!r})"

    def starts_with(self, prefix: "Kmer") -> bool:
        """`kmer ^@ prefix` for ONE value (dna.c:842-866); batches go through the GPU."""
        if prefix.length > self.length:
            raise DnaError("Prefix length cannot exceed kmer length")  # dna.c:854-856
        mask = (1 << (2 * prefix.length)) - 1  # 32 -> full 64 bits (Q1)
        return prefix.bits == (self.bits & mask)


class Qkmer:
    """A `qkmer` value: an IUPAC pattern of 1..32 characters (dna.c:81-84, 876-900)."""

    __slots__ = ("pattern",)

    def __init__(self, text):
        if isinstance(text, Qkmer):
            self.pattern = text.pattern
            return
        if text == "":
            raise DnaError("qkmer pattern cannot be empty")  # dna.c:877-879
        if len(text) > 32:
            raise DnaError("Qkmer pattern length cannot exceed 32 characters")  # dna.c:883-885
        for ch in text:
            if ch not in QKMER_ALPHABET:
                raise DnaError(f"Invalid character in qkmer pattern: {ch}")  # dna.c:893-895
        self.pattern = text

    def __len__(self):
        return len(self.pattern)

    def __str__(self):
        return self.pattern

    def __eq__(self, other):
        return isinstance(other, Qkmer) and self.pattern == other.pattern

    def __repr__(self):
        return f"Qkmer({self.pattern!r})"


def kmer_strings(bits: np.ndarray, k: int):
    """Decode an array of kmer bit_sequences to text (kmer_out), vectorised."""
    bits = np.ascontiguousarray(bits, dtype=np.uint64)
    out = np.empty((bits.size, k), dtype=np.uint8)
    for j in range(k):
        out[:, j] = _LETTERS[((bits >> np.uint64(2 * j)) & np.uint64(3)).astype(np.uint8)]
    return [row.tobytes().decode() for row in out]
