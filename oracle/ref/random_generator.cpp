// See random_generator.h in this directory; default path mirrors
// /root/reference/src/random_generator.cpp:4-11.
#include "random_generator.h"

#include <limits>

static thread_local const float *t_replay = nullptr;
static thread_local int t_replayCount = 0;
static thread_local int t_replayUsed = 0;

RandomGenerator::RandomGenerator()
    : m_generator(m_device()), m_distribution(0.f, 1.f - std::numeric_limits<float>::epsilon())
{}

float RandomGenerator::next()
{
    if (t_replay) {
        const float xi = t_replay[t_replayUsed % t_replayCount];
        t_replayUsed += 1;
        return xi;
    }
    return m_distribution(m_generator);
}

void RandomGenerator::beginReplay(const float *xi, int count)
{
    t_replay = xi;
    t_replayCount = count;
    t_replayUsed = 0;
}

int RandomGenerator::endReplay()
{
    const int used = t_replayUsed;
    t_replay = nullptr;
    t_replayCount = 0;
    t_replayUsed = 0;
    return used;
}
