"""splendor-ai_b200/actions.py (closed-form ALL_ACTIONS) vs the dump of the reference's ALL_ACTIONS list
(tests/golden/tables.json, from gym/envs/actions.py:222-237)."""
import json
import os

from conftest import GOLDEN

TABLES = json.load(open(os.path.join(GOLDEN, "tables.json")))


def test_decode_matches_reference_all_actions():
    from splendor_ai_b200 import actions as A

    assert len(A.ALL_ACTIONS) == len(TABLES["all_actions"]) == 3510
    for idx, (tname, collected, returned, pos, noble) in enumerate(TABLES["all_actions"]):
        a = A.ALL_ACTIONS[idx]
        assert a.type_enum.name == tname, idx
        assert a.collected_gems == collected and a.returned_gems == returned, idx
        assert a.noble_index == noble, idx
        assert (None if a.position is None else [a.position.tier, a.position.card_index, a.position.reserved_index]) == pos, idx


def test_index_is_the_inverse_of_decode():
    from splendor_ai_b200 import actions as A

    for idx in range(0, 3510):
        assert A.ALL_ACTIONS.index(A.decode(idx)) == idx


def test_out_of_range_raises_value_error():
    import pytest

    from splendor_ai_b200 import actions as A

    for bad in (-1, 3510, 10**6):
        with pytest.raises(ValueError):
            A.decode(bad)
