import re


class RegexpTokenizer(object):
    """The one nltk name datasets.py imports (datasets.py:6); same call contract for plain patterns."""

    def __init__(self, pattern):
        self._re = re.compile(pattern)

    def tokenize(self, text):
        return self._re.findall(text)
