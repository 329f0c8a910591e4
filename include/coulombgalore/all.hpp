// coulombgalore/all.hpp -- umbrella include (reference all.hpp:27-38) and the run-time factory createScheme
// (reference all.hpp:42-104).  The reference's factory reads a nlohmann::json object; this one takes the same content
// without the dependency: the "type" keyword and a map of the numeric keys (cutoff, alpha, order, C, D, epsRF, epsr,
// shifted, epss, debyelength).  Same behaviour: unknown keyword -> std::runtime_error("unknown coulomb scheme ..."),
// missing mandatory key -> std::out_of_range (json::at), "spline" -> null pointer.
#pragma once
#include "batched.hpp"
#include "ewald.hpp"
#include "ewald_truncated.hpp"
#include "fanourgakis.hpp"
#include "fennell.hpp"
#include "plain.hpp"
#include "poisson.hpp"
#include "qpotential.hpp"
#include "reactionfield.hpp"
#include "reciprocal.hpp"
#include "splined.hpp"
#include "wolf.hpp"
#include "zahn.hpp"
#include "zerodipole.hpp"

#include <map>
#include <memory>
#include <stdexcept>
#include <string>

namespace CoulombGalore {

inline std::shared_ptr<SchemeBase> createScheme(const std::string &type, const std::map<std::string, double> &j) {
    auto at = [&](const char *key) {
        const auto it = j.find(key);
        if (it == j.end()) throw std::out_of_range(std::string("key '") + key + "' not found");
        return it->second;
    };
    auto value = [&](const char *key, double dflt) {
        const auto it = j.find(key);
        return it == j.end() ? dflt : it->second;
    };
    if (type == "plain") return std::make_shared<Plain>(value("debyelength", infinity));
    if (type == "wolf") return std::make_shared<Wolf>(at("cutoff"), at("alpha"));
    if (type == "zahn") return std::make_shared<Zahn>(at("cutoff"), at("alpha"));
    if (type == "fennell") return std::make_shared<Fennell>(at("cutoff"), at("alpha"));
    if (type == "zerodipole") return std::make_shared<ZeroDipole>(at("cutoff"), at("alpha"));
    if (type == "fanourgakis") return std::make_shared<Fanourgakis>(at("cutoff"));
    if (type == "qpotential") return std::make_shared<qPotential>(at("cutoff"), (int)at("order"));
    if (type == "ewald") return std::make_shared<Ewald>(at("cutoff"), at("alpha"), value("epss", infinity), value("debyelength", infinity));
    if (type == "ewaldt")
        return std::make_shared<EwaldTruncated>(at("cutoff"), at("alpha"), value("epss", infinity), value("debyelength", infinity));
    if (type == "poisson") return std::make_shared<Poisson>(at("cutoff"), (int)at("C"), (int)at("D"), value("debyelength", infinity));
    if (type == "reactionfield") return std::make_shared<ReactionField>(at("cutoff"), at("epsRF"), at("epsr"), at("shifted") != 0.0);
    if (type == "spline") return nullptr; // a known keyword the reference's switch has no case for (all.hpp:98-100)
    throw std::runtime_error("unknown coulomb scheme " + type);
}

} // namespace CoulombGalore
