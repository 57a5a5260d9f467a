"""Rule names and orders of the reference's data.yaml, restated as Python constants.

Orders are semantics: inventory / item-strip order (data.yaml:39-55, engine.py:230,238), achievement
order (data.yaml:80-102), action ids (data.yaml:1-18), material ids = index + 1 (engine.py:29-30).
The numeric rule tables themselves (collect / place / make) live in csrc/cr_update.h.
"""

ACTIONS = [
    'noop', 'move_left', 'move_right', 'move_up', 'move_down', 'do', 'sleep', 'place_stone',
    'place_table', 'place_furnace', 'place_plant', 'make_wood_pickaxe', 'make_stone_pickaxe',
    'make_iron_pickaxe', 'make_wood_sword', 'make_stone_sword', 'make_iron_sword']

MATERIALS = [
    'water', 'grass', 'stone', 'path', 'sand', 'tree', 'lava', 'coal', 'iron', 'diamond', 'table',
    'furnace']

ITEMS = [
    'health', 'food', 'drink', 'energy', 'sapling', 'wood', 'stone', 'coal', 'iron', 'diamond',
    'wood_pickaxe', 'stone_pickaxe', 'iron_pickaxe', 'wood_sword', 'stone_sword', 'iron_sword']

ACHIEVEMENTS = [
    'collect_coal', 'collect_diamond', 'collect_drink', 'collect_iron', 'collect_sapling',
    'collect_stone', 'collect_wood', 'defeat_skeleton', 'defeat_zombie', 'eat_cow', 'eat_plant',
    'make_iron_pickaxe', 'make_iron_sword', 'make_stone_pickaxe', 'make_stone_sword',
    'make_wood_pickaxe', 'make_wood_sword', 'place_furnace', 'place_plant', 'place_stone',
    'place_table', 'wake_up']

# Sprite order of the object atlas (csrc/cr_common.h ObjTex).
OBJECT_SPRITES = [
    'player-left', 'player-right', 'player-up', 'player-down', 'player-sleep', 'cow', 'zombie',
    'skeleton', 'arrow-left', 'arrow-right', 'arrow-up', 'arrow-down', 'plant', 'plant-ripe']

# Semantic-view ids (engine.py:253-258 with env.py:47-49): materials 1..12, then these.
SEMANTIC_OBJECTS = ['player', 'cow', 'zombie', 'skeleton', 'arrow', 'plant']  # ids 13..18

# csrc/cr_common.h PState columns.
PSTATE = ['hunger2', 'thirst2', 'fatigue', 'recover2', 'sleeping', 'player_last_health',
          'env_last_health', 'unlocked', 'n_slots', 'step', 'episode', 'world_seed', 'player_x',
          'player_y', 'error', 'episode_length']
