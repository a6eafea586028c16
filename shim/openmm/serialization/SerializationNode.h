#ifndef SHIM_SERIALIZATION_NODE_H_
#define SHIM_SERIALIZATION_NODE_H_
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include "openmm/OpenMMException.h"
namespace OpenMM {
/** Typed string properties + children, like OpenMM's SerializationNode (doubles keep 17 significant digits). */
class SerializationNode {
public:
    const std::string& getName() const { return name; }
    void setName(const std::string& n) { name = n; }
    const std::vector<SerializationNode>& getChildren() const { return children; }
    std::vector<SerializationNode>& getChildren() { return children; }
    const std::map<std::string, std::string>& getProperties() const { return properties; }
    bool hasProperty(const std::string& n) const { return properties.count(n) != 0; }
    const std::string& getStringProperty(const std::string& n) const {
        std::map<std::string, std::string>::const_iterator it = properties.find(n);
        if (it == properties.end()) throw OpenMMException("Unknown property '" + n + "' in node '" + name + "'");
        return it->second;
    }
    SerializationNode& setStringProperty(const std::string& n, const std::string& v) { properties[n] = v; return *this; }
    int getIntProperty(const std::string& n) const { int v; std::stringstream(getStringProperty(n)) >> v; return v; }
    int getIntProperty(const std::string& n, int d) const { return hasProperty(n) ? getIntProperty(n) : d; }
    SerializationNode& setIntProperty(const std::string& n, int v) { std::stringstream s; s << v; properties[n] = s.str(); return *this; }
    bool getBoolProperty(const std::string& n) const { return getIntProperty(n) != 0; }
    bool getBoolProperty(const std::string& n, bool d) const { return hasProperty(n) ? getBoolProperty(n) : d; }
    SerializationNode& setBoolProperty(const std::string& n, bool v) { return setIntProperty(n, v ? 1 : 0); }
    double getDoubleProperty(const std::string& n) const { double v; std::stringstream(getStringProperty(n)) >> v; return v; }
    double getDoubleProperty(const std::string& n, double d) const { return hasProperty(n) ? getDoubleProperty(n) : d; }
    SerializationNode& setDoubleProperty(const std::string& n, double v) { std::stringstream s; s.precision(17); s << v; properties[n] = s.str(); return *this; }
    SerializationNode& createChildNode(const std::string& n) { children.push_back(SerializationNode()); children.back().setName(n); return children.back(); }
    const SerializationNode& getChildNode(const std::string& n) const {
        for (size_t i = 0; i < children.size(); i++) if (children[i].name == n) return children[i];
        throw OpenMMException("Unknown child node '" + n + "' in node '" + name + "'");
    }
private:
    std::string name;
    std::vector<SerializationNode> children;
    std::map<std::string, std::string> properties;
};
}  // namespace OpenMM
#endif
