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
