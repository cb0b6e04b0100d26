// testing_b200.cu -- the reference's test-driver sequence (testing.cu:51-110: Simulation(1024, 100),
// CPU pricing from pre-generated normals, the four reductions, outer trajectories -> testing.csv)
// written against the compat headers, plus the GPU twin of the pre-generated-normal pricer and the
// library's own CSV writer.  The reference's testing.cu itself compiles unchanged the same way.
#include "testing.cuh"

#include <fstream>

int main(int argc, char **argv)
{
    const char *csv = argc > 1 ? argv[1] : "testing.csv";
    Simulation sim(1024, 100);

    auto cpu = sim.simulate_trajectory_cpu();
    auto gpu = sim.simulate_trajectory_gpu();
    double worst = 0.0;
    for (size_t i = 0; i < cpu.size(); ++i) worst = fmax(worst, fabs((double)cpu[i] - (double)gpu[i]));
    printf("PREGEN %zu %.9g\n", cpu.size(), worst);

    const float host_sum = sim.sum_random_array();
    for (int kind = SequentialAddressing; kind <= CompletelyUnrolled; ++kind) {
        auto out = sim.test_reduction(1, 1024, kind);
        printf("REDUCE %d %.9g\n", kind, out[0]);
    }
    printf("HOSTSUM %.9g\n", host_sum);

    Simulation outer(20, 150);
    auto rows = outer.simulate_outer_trajectories(10, 555);
    printf("OUTER %zu %.9g %.9g\n", rows.size(), rows.front(), rows.back());
    if (mcb_write_trajectories_csv(csv, rows.data(), outer.n_trajectories, (int)outer.n_steps, outer.x_0, outer.dt()) != MCB_OK) {
        fprintf(stderr, "%s\n", mcb_last_error());
        return 1;
    }
    // the same file written the way testing.cu:37-47 does, to prove the formats agree byte for byte
    std::ofstream ref(std::string(csv) + ".ostream");
    ref << "time,trajectory,value\n";
    const int n_traj = 20, n_steps = 150;
    for (int i = 0; i < n_traj * n_steps; i++) {
        int i_traj = i / n_steps;
        if (i % n_steps == 0) ref << 0.0 << "," << i_traj << "," << outer.x_0 << "\n";
        ref << (1 + i % n_steps) * outer.dt() << "," << i_traj << "," << rows[i] << "\n";
    }
    return 0;
}
