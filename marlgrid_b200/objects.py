"""Object-type / colour tables of the world model.

Mirrors the numeric encodings of the reference so encoded observations are drop-in:
  * type index  = position in OBJECT_TYPES, filled by the RegisteredObjectType metaclass in class
    definition order (marlgrid/objects.py:31-43) plus GridAgentInterface (marlgrid/agents.py:9)
  * colour index = key order of COLORS (marlgrid/objects.py:11-29)
  * Door states  (marlgrid/objects.py:325)
The per-type behaviour predicates (can_overlap / can_pickup / see_behind, objects.py:75-88,
147-148,174,216,230,258,281,292,314,327-331,378) live in the CUDA kernels as constant tables
(marlgrid_b200/csrc/mg_tables.cuh); the Python copies below are for host-side inspection.
"""
import enum

import numpy as np

OBJECT_TYPE_NAMES = (
    "WorldObj", "GridAgent", "BulkObj", "BonusTile", "Goal", "Floor", "EmptySpace", "Lava",
    "Wall", "Key", "Ball", "Door", "Box", "GridAgentInterface",
)
TYPE_TO_IDX = {n: i for i, n in enumerate(OBJECT_TYPE_NAMES)}
T_EMPTY, T_BONUS, T_GOAL, T_FLOOR, T_LAVA, T_WALL, T_KEY, T_BALL, T_DOOR, T_BOX, T_AGENT = 0, 3, 4, 5, 7, 8, 9, 10, 11, 12, 13

# Map of colour names to RGB values (same table as marlgrid/objects.py:11-26)
COLORS = {
    "red": np.array([255, 0, 0]),
    "orange": np.array([255, 165, 0]),
    "green": np.array([0, 255, 0]),
    "blue": np.array([0, 0, 255]),
    "cyan": np.array([0, 139, 139]),
    "purple": np.array([112, 39, 195]),
    "yellow": np.array([255, 255, 0]),
    "olive": np.array([128, 128, 0]),
    "grey": np.array([100, 100, 100]),
    "worst": np.array([74, 65, 42]),
    "pink": np.array([255, 0, 189]),
    "white": np.array([255, 255, 255]),
    "prestige": np.array([255, 255, 255]),
    "shadow": np.array([35, 25, 30]),
}
COLOR_TO_IDX = {k: i for i, k in enumerate(COLORS.keys())}
IDX_TO_COLOR = {i: k for k, i in COLOR_TO_IDX.items()}

DOOR_OPEN, DOOR_CLOSED, DOOR_LOCKED = 1, 2, 3



class Actions(enum.IntEnum):
    """GridAgentInterface.actions (marlgrid/agents.py:10-17): `agent.actions.forward`, `len(agent.actions)`."""
    left = 0      # Rotate left
    right = 1     # Rotate right
    forward = 2   # Move forward
    pickup = 3    # Pick up an object
    drop = 4      # Drop an object
    toggle = 5    # Toggle/activate an object
    done = 6      # Done completing task


ACTIONS = {a.name: int(a) for a in Actions}


def can_overlap(type_idx, state=0):
    return type_idx in (T_BONUS, T_GOAL, T_FLOOR, T_LAVA, T_AGENT, 1) or (type_idx == T_DOOR and state == DOOR_OPEN)


def can_pickup(type_idx):
    return type_idx in (T_KEY, T_BALL, T_BOX)


def see_behind(type_idx, state=0):
    if type_idx == T_WALL:
        return False
    if type_idx == T_DOOR:
        return state == DOOR_OPEN
    return True


def hide_mask(type_names):
    """GridAgentInterface.hide_item_types (a list of WorldObj.type strings, agents.py:30; 'Agent' for agents,
    objects.py:137-138) -> the MgConfig.hide_types bit set over type indices."""
    mask = 0
    for name in type_names:
        if name == "Agent":
            mask |= 1 << T_AGENT
        elif name in TYPE_TO_IDX:
            mask |= 1 << TYPE_TO_IDX[name]
        else:
            raise ValueError(f"unknown object type {name!r} in hide_item_types")
    return mask
