def is_configured():
    return False
