def pipeline(task, model=None):
    """Returns an 'OCR' callable that always reads five in-alphabet characters."""
    def run(img):
        return {"text": ["ABCDE"]}
    return run
