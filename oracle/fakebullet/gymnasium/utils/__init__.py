"""TEST INFRASTRUCTURE stub."""


def colorize(string, *a, **k):
    return string
