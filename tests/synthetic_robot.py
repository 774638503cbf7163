"""A synthetic 9-DOF arm URDF (mixed joint axes, rpy origins, one continuous joint, a welded link mid-chain, a frozen
finger joint, 2 spheres per link) used to exercise the general robot compiler and the A=16/32, NSLOT=4 kernel
instantiations.  Test infrastructure only."""
import numpy as np

AXES = ["0 0 1", "0 1 0", "1 0 0", "0 1 0", "0 0 1", "0 1 0", "1 0 0", "0 0 1", "0 1 0"]


def urdf_text(n_joints: int = 9, spheres_per_link: int = 2, seed: int = 0) -> str:
    rng = np.random.RandomState(seed)
    out = ['<?xml version="1.0"?>', '<robot name="synth9">', '<link name="base"><collision><origin xyz="0 0 0.05"/>'
           '<geometry><box size="0.3 0.3 0.1"/></geometry></collision></link>']
    parent = "base"
    for j in range(n_joints):
        link = f"l{j}"
        kind = "continuous" if j == 4 else "revolute"
        xyz = [0.0, 0.0, 0.16] if j else [0.0, 0.0, 0.12]
        rpy = rng.uniform(-0.4, 0.4, 3) * (j % 2)
        out.append(f'<joint name="j{j}" type="{kind}"><parent link="{parent}"/><child link="{link}"/>'
                   f'<origin xyz="{xyz[0]} {xyz[1]} {xyz[2]}" rpy="{rpy[0]:.6f} {rpy[1]:.6f} {rpy[2]:.6f}"/>'
                   f'<axis xyz="{AXES[j % len(AXES)]}"/>' + ('' if kind == "continuous" else
                   f'<limit lower="{-2.0 - 0.1 * j:.3f}" upper="{2.0 + 0.05 * j:.3f}" effort="10" velocity="1"/>') + '</joint>')
        cols = ""
        for k in range(spheres_per_link):
            z = 0.04 + 0.08 * k
            r = 0.035 + 0.004 * ((j + k) % 3)
            cols += (f'<collision><origin xyz="{0.01 * (k - 0.5):.4f} 0 {z:.4f}"/><geometry><sphere radius="{r:.4f}"/>'
                     f'</geometry></collision>')
        out.append(f'<link name="{link}"><inertial><origin xyz="0 0.01 0.05"/><mass value="1"/></inertial>{cols}</link>')
        parent = link
        if j == 5:      # welded sensor link in the middle of the chain, carrying a sphere
            out.append(f'<joint name="weld" type="fixed"><parent link="{parent}"/><child link="sensor"/>'
                       f'<origin xyz="0.05 0 0.02" rpy="0 0.3 0"/></joint>')
            out.append('<link name="sensor"><collision><origin xyz="0 0 0.02"/><geometry><sphere radius="0.02"/></geometry>'
                       '</collision></link>')
    out.append(f'<joint name="finger" type="revolute"><parent link="{parent}"/><child link="tip"/>'
               f'<origin xyz="0 0.02 0.1" rpy="0 0 0"/><axis xyz="1 0 0"/><limit lower="0" upper="1" effort="1" velocity="1"/></joint>')
    out.append('<link name="tip"><inertial><origin xyz="0 0 0.03"/><mass value="0.1"/></inertial>'
               '<collision><origin xyz="0 0 0.03"/><geometry><sphere radius="0.015"/></geometry></collision></link>')
    out.append('</robot>')
    return "\n".join(out)


def synthetic_robot(n_joints: int = 9, spheres_per_link: int = 2, seed: int = 0, prune: bool = True):
    from plant_motion_planning_b200.model.robot import compile_robot, prune_always_colliding
    from plant_motion_planning_b200.model.urdf import parse_urdf
    tree = parse_urdf(urdf_text(n_joints, spheres_per_link, seed), is_text=True)
    model = compile_robot(tree, active_joints=[f"j{j}" for j in range(n_joints)], frozen_positions={"finger": 0.4},
                          ee_link="tip", name="synth9")
    if prune:
        prune_always_colliding(model, samples=512)
    return model


def val_like_urdf_text(seed: int = 0) -> str:
    """A robot with Val's joint layout (pybullet_tools/val_utils.py:5-23): movable joints 0-3 wheels, 4-5 torso,
    6-14 the nine left-arm joints -- so that the reference's BodyConf / get_arm_joints pick the arm unchanged."""
    rng = np.random.RandomState(seed)
    out = ['<?xml version="1.0"?>', '<robot name="val_like">',
           '<link name="base"><collision><origin xyz="0 0 0.1"/><geometry><box size="0.5 0.4 0.2"/></geometry></collision></link>']
    for k, (x, y) in enumerate([(0.2, 0.2), (0.2, -0.2), (-0.2, 0.2), (-0.2, -0.2)]):
        out.append(f'<joint name="wheel{k}" type="continuous"><parent link="base"/><child link="w{k}"/>'
                   f'<origin xyz="{x} {y} 0.05"/><axis xyz="0 1 0"/></joint><link name="w{k}"/>')
    out.append('<joint name="torso0" type="revolute"><parent link="base"/><child link="t0"/><origin xyz="0 0 0.25"/>'
               '<axis xyz="0 0 1"/><limit lower="-1" upper="1" effort="1" velocity="1"/></joint>'
               '<link name="t0"><collision><origin xyz="0 0 0.1"/><geometry><sphere radius="0.09"/></geometry></collision></link>')
    out.append('<joint name="torso1" type="revolute"><parent link="t0"/><child link="t1"/><origin xyz="0 0 0.2"/>'
               '<axis xyz="0 1 0"/><limit lower="-1" upper="1" effort="1" velocity="1"/></joint>'
               '<link name="t1"><collision><origin xyz="0 0 0.1"/><geometry><sphere radius="0.08"/></geometry></collision></link>')
    parent = "t1"
    for j in range(9):
        link = f"a{j}"
        xyz = [0.0, 0.12, 0.15] if j == 0 else [0.0, 0.0, 0.11]
        rpy = rng.uniform(-0.5, 0.5, 3) * (j % 2)
        out.append(f'<joint name="arm{j}" type="revolute"><parent link="{parent}"/><child link="{link}"/>'
                   f'<origin xyz="{xyz[0]} {xyz[1]} {xyz[2]}" rpy="{rpy[0]:.6f} {rpy[1]:.6f} {rpy[2]:.6f}"/>'
                   f'<axis xyz="{AXES[j]}"/><limit lower="{-2.2 - 0.05 * j:.3f}" upper="{2.2 + 0.05 * j:.3f}" effort="10" velocity="1"/></joint>')
        r = 0.04 - 0.002 * j
        out.append(f'<link name="{link}"><inertial><origin xyz="0 0 0.05"/><mass value="1"/></inertial>'
                   f'<collision><origin xyz="0 0 0.055"/><geometry><sphere radius="{r:.4f}"/></geometry></collision></link>')
        parent = link
    out.append('</robot>')
    return "\n".join(out)


def val_like_robot(seed: int = 0, torso=(0.2, -0.3), ee_link_index: int = 9):
    """Compiled with the arm as the active chain, wheels/torso frozen; the cost end-effector is PyBullet link 9, the
    link the reference hard-codes (kuka_primitives.py:478)."""
    from plant_motion_planning_b200.model.robot import compile_robot, prune_always_colliding
    from plant_motion_planning_b200.model.transforms import Frame
    from plant_motion_planning_b200.model.urdf import parse_urdf
    tree = parse_urdf(val_like_urdf_text(seed), is_text=True)
    model = compile_robot(tree, active_joints=[f"arm{j}" for j in range(9)],
                          frozen_positions={"torso0": torso[0], "torso1": torso[1]},
                          base=Frame(np.eye(3), np.array([-0.35, 0.25, 0.0])),
                          ee_link=tree.joints[ee_link_index].child, name="val_like")
    prune_always_colliding(model, samples=512)
    return model, tree
