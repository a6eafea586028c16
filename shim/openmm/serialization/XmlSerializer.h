#ifndef SHIM_XML_SERIALIZER_H_
#define SHIM_XML_SERIALIZER_H_
#include <iostream>
#include <sstream>
#include <string>
#include <typeinfo>
#include "openmm/serialization/SerializationNode.h"
#include "openmm/serialization/SerializationProxy.h"
namespace OpenMM {
/** XML in the shape OpenMM writes: <Root type="..." version="..." prop="..."> children </Root>. */
class XmlSerializer {
public:
    template <class T>
    static void serialize(const T* object, const std::string& rootName, std::ostream& stream) {
        const SerializationProxy& proxy = SerializationProxy::getProxy(typeid(*object));
        SerializationNode node;
        node.setName(rootName);
        proxy.serialize(object, node);
        node.setStringProperty("type", proxy.getTypeName());
        stream << "<?xml version=\"1.0\" ?>\n";
        write(node, stream, 0);
    }
    template <class T>
    static T* deserialize(std::istream& stream) {
        std::stringstream all;
        all << stream.rdbuf();
        std::string text = all.str();
        size_t pos = text.find("?>");
        pos = (pos == std::string::npos) ? 0 : pos + 2;
        SerializationNode node = parse(text, pos);
        const SerializationProxy& proxy = SerializationProxy::getProxy(node.getStringProperty("type"));
        return reinterpret_cast<T*>(proxy.deserialize(node));
    }
private:
    static void write(const SerializationNode& node, std::ostream& out, int depth) {
        out << std::string(depth, '\t') << '<' << node.getName();
        for (std::map<std::string, std::string>::const_iterator it = node.getProperties().begin(); it != node.getProperties().end(); ++it)
            out << ' ' << it->first << "=\"" << it->second << '"';
        if (node.getChildren().empty()) { out << "/>\n"; return; }
        out << ">\n";
        for (size_t i = 0; i < node.getChildren().size(); i++) write(node.getChildren()[i], out, depth + 1);
        out << std::string(depth, '\t') << "</" << node.getName() << ">\n";
    }
    static void skipSpace(const std::string& t, size_t& p) { while (p < t.size() && isspace((unsigned char)t[p])) p++; }
    static SerializationNode parse(const std::string& t, size_t& p) {
        skipSpace(t, p);
        if (p >= t.size() || t[p] != '<') throw OpenMMException("XML parse error: expected '<'");
        p++;
        size_t e = p;
        while (e < t.size() && !isspace((unsigned char)t[e]) && t[e] != '>' && t[e] != '/') e++;
        SerializationNode node;
        node.setName(t.substr(p, e - p));
        p = e;
        while (true) {
            skipSpace(t, p);
            if (t.compare(p, 2, "/>") == 0) { p += 2; return node; }
            if (t[p] == '>') { p++; break; }
            size_t eq = t.find('=', p);
            std::string key = t.substr(p, eq - p);
            size_t q1 = t.find('"', eq), q2 = t.find('"', q1 + 1);
            node.setStringProperty(key, t.substr(q1 + 1, q2 - q1 - 1));
            p = q2 + 1;
        }
        while (true) {
            skipSpace(t, p);
            if (t.compare(p, 2, "</") == 0) { p = t.find('>', p) + 1; return node; }
            node.getChildren().push_back(parse(t, p));
        }
    }
};
}  // namespace OpenMM
#endif
