"""Host-side mirror of the reference's ``models`` package (networks / resnet / textvit_arch / ocr)."""
