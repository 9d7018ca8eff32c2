// Test-infrastructure replacement for the reference's include/random_generator.h
// (/root/reference/include/random_generator.h:5-15).  Same class, same members, same
// default behaviour (mt19937 seeded from random_device, uniform [0, 1-eps)); adds a
// thread-local replay queue so the probe harness can feed explicit xi values to
// Material::sample / Light::sample and compare them with the CUDA path on equal inputs.
#pragma once

#include <random>

class RandomGenerator {
public:
    RandomGenerator();

    float next();

    // probe-only: while a replay buffer is installed on this thread, next() pops from it
    static void beginReplay(const float *xi, int count);
    static int endReplay();   // returns how many values were consumed

private:
    std::random_device m_device;
    std::mt19937 m_generator;
    std::uniform_real_distribution<float> m_distribution;
};
