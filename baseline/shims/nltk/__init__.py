"""Import stand-in for nltk (absent from this image): the reference's datasets.py:6 imports RegexpTokenizer at module level;
the eval flow under test never tokenises (captions arrive as indices)."""
