"""Stand-in for modelscope (third-party OCR pipeline, not installed). TEST INFRASTRUCTURE ONLY."""
