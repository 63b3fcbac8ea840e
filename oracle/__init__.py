"""Test infrastructure: CPU oracle of the WS-MGMap map update.  Never imported by the product path."""
