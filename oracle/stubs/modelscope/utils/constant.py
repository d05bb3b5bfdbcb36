class Tasks:
    ocr_recognition = "ocr-recognition"
