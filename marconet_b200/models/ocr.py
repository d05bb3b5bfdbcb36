"""Import-compatibility module.  The reference's ``models/ocr.py`` (legacy TransformerOCR) is imported by
test_sr.py:6 but never instantiated on the inference path; only the name needs to exist."""


class TransformerOCR:  # pragma: no cover - dead code on the reference path
    def __init__(self, *a, **k):
        raise NotImplementedError("TransformerOCR is not part of the MARCONet inference hot path (out of scope)")
