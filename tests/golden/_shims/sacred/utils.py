def apply_backspaces_and_linefeeds(text):
    return text
