"""CLIP byte-pair tokenizer (decode side is what the captioner needs: `open_clip.decode(tokens)` at
`experimenting_env/captioner/models/coca/coca.py:30`, `scripts/evaluate_finetuned_model.py:44`).

Restates upstream open_clip `SimpleTokenizer`: 256 byte symbols + 256 `</w>` variants + 48 894 merges +
`<start_of_text>` / `<end_of_text>` = 49 408 entries.  The merge table
(`bpe_simple_vocab_16e6.txt.gz`) ships with open_clip, which is not available offline; point
`EC_BPE_VOCAB` (or `bpe_path=`) at it.  Without it, `decode` renders ids symbolically (`<tok_123>`) so
that the rest of the pipeline (string post-processing at coca.py:30) still runs; that is a rendering
fallback only -- token ids, which carry the parity contract, are unaffected.
"""
from __future__ import annotations

import gzip
import os
from functools import lru_cache
from typing import Iterable, List, Optional

SOT_TEXT = "<start_of_text>"
EOT_TEXT = "<end_of_text>"


@lru_cache()
def bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(2 ** 8):
        if b not in bs:
            bs.append(b)
            cs.append(2 ** 8 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def _find_vocab(explicit: Optional[str]) -> Optional[str]:
    cands = [explicit, os.environ.get("EC_BPE_VOCAB"),
             os.path.join(os.path.dirname(__file__), "bpe_simple_vocab_16e6.txt.gz")]
    for c in cands:
        if c and os.path.isfile(c):
            return c
    return None


class SimpleTokenizer:
    def __init__(self, bpe_path: Optional[str] = None, vocab_size: int = 49408, context_length: int = 77):
        self.context_length = context_length
        self.vocab_size = vocab_size
        self.sot_token_id = vocab_size - 2
        self.eot_token_id = vocab_size - 1
        self.byte_encoder = bytes_to_unicode()
        self.byte_decoder = {v: k for k, v in self.byte_encoder.items()}
        path = _find_vocab(bpe_path)
        self.has_vocab = path is not None
        self.decoder = {}
        self._cache = {SOT_TEXT: [SOT_TEXT], EOT_TEXT: [EOT_TEXT]}
        self._pat = None
        if self.has_vocab:
            merges = gzip.open(path).read().decode("utf-8").split("\n")
            # upstream keeps `49152 - 256 - 2 + 1` lines = 48 894 merges after the header: 512 byte symbols +
            # merges + 2 specials = 49 408, so <start_of_text> / <end_of_text> are the last two ids
            merges = merges[1:vocab_size - 512 - 2 + 1]
            merges = [tuple(m.split()) for m in merges if m.strip()]
            vocab = list(self.byte_encoder.values())
            vocab = vocab + [v + "</w>" for v in vocab]
            vocab.extend("".join(m) for m in merges)
            vocab.extend([SOT_TEXT, EOT_TEXT])
            if len(vocab) != vocab_size:
                raise RuntimeError(f"BPE table {path} yields {len(vocab)} entries, the model has {vocab_size} "
                                   "(merge file too short for this vocabulary)")
            self.encoder = dict(zip(vocab, range(len(vocab))))
            assert self.encoder[SOT_TEXT] == self.sot_token_id and self.encoder[EOT_TEXT] == self.eot_token_id
            self.decoder = {v: k for k, v in self.encoder.items()}
            self.bpe_ranks = dict(zip(merges, range(len(merges))))

    def decode(self, tokens: Iterable[int]) -> str:
        toks: List[int] = [int(t) for t in (tokens.tolist() if hasattr(tokens, "tolist") else tokens)]
        if not self.has_vocab:
            parts = []
            for t in toks:
                if t == self.sot_token_id:
                    parts.append(SOT_TEXT)
                elif t == self.eot_token_id:
                    parts.append(EOT_TEXT)
                else:
                    parts.append(f"<tok_{t}> ")
            return "".join(parts)
        text = "".join(self.decoder[t] for t in toks)
        return bytearray([self.byte_decoder[c] for c in text]).decode("utf-8", errors="replace").replace("</w>", " ")

    # ------------------------------------------------------------------ encode side (CLIP byte-level BPE)
    def _merge_word(self, word: str) -> List[str]:
        """Greedy lowest-rank-first pair merging of one pre-token; the last symbol carries the `</w>` marker."""
        cached = self._cache.get(word)
        if cached is not None:
            return cached
        syms = list(word[:-1]) + [word[-1] + "</w>"]
        while len(syms) > 1:
            best_rank, best_at = None, -1
            for i in range(len(syms) - 1):
                r = self.bpe_ranks.get((syms[i], syms[i + 1]))
                if r is not None and (best_rank is None or r < best_rank):
                    best_rank, best_at = r, i
            if best_rank is None:
                break
            a, b = syms[best_at], syms[best_at + 1]
            out, i = [], 0
            while i < len(syms):                       # merge EVERY occurrence of the winning pair, left to right
                if i + 1 < len(syms) and syms[i] == a and syms[i + 1] == b:
                    out.append(a + b); i += 2
                else:
                    out.append(syms[i]); i += 1
            syms = out
        self._cache[word] = syms
        return syms

    def encode(self, text: str) -> List[int]:
        if not self.has_vocab:
            raise RuntimeError("encode() needs the BPE merge table (EC_BPE_VOCAB / bpe_path=)")
        import html
        import regex
        if self._pat is None:
            self._pat = regex.compile(r"<start_of_text>|<end_of_text>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|"
                                      r"[^\s\p{L}\p{N}]+", regex.IGNORECASE)
        text = html.unescape(html.unescape(text))
        text = " ".join(text.split()).strip().lower()
        ids: List[int] = []
        for piece in self._pat.findall(text):
            if piece in (SOT_TEXT, EOT_TEXT):
                ids.append(self.encoder[piece])
                continue
            mapped = "".join(self.byte_encoder[b] for b in piece.encode("utf-8"))
            ids.extend(self.encoder[sym] for sym in self._merge_word(mapped))
        return ids

    def __call__(self, texts, context_length: Optional[int] = None):
        """`open_clip.tokenize`: int64 [N, context_length] = SOT + ids + EOT, zero padded; over-long rows are cut and
        end with EOT."""
        import torch
        if isinstance(texts, str):
            texts = [texts]
        L = context_length or self.context_length
        out = torch.zeros(len(texts), L, dtype=torch.long)
        for i, t in enumerate(texts):
            ids = [self.sot_token_id] + self.encode(t) + [self.eot_token_id]
            if len(ids) > L:
                ids = ids[:L]
                ids[-1] = self.eot_token_id
            out[i, :len(ids)] = torch.tensor(ids)
        return out


_default: Optional[SimpleTokenizer] = None


def get_tokenizer(model_name: str = "", vocab_size: int = 49408, **kw) -> SimpleTokenizer:
    return SimpleTokenizer(vocab_size=vocab_size, **kw)


def decode(tokens, vocab_size: int = 49408) -> str:
    """Module-level `open_clip.decode` equivalent."""
    global _default
    if _default is None or _default.vocab_size != vocab_size:
        _default = SimpleTokenizer(vocab_size=vocab_size)
    return _default.decode(tokens)
